"""Known-answer vectors held by the reference's own unit tests for the hot path.

Each entry cites the reference test it comes from; patterns are re-generated from the
formulas in those tests (nothing is loaded from /root/reference at run time).
"""
import math
import numpy as np

# tests/satd_tests.c:122,140,159 - satd_{4,8,16,32,64} on three patterns
SATD_BLACK_WHITE = [510, 1020, 4080, 16320, 65280]
SATD_CHECKERS = [1278, 2556, 10224, 40896, 163584]
SATD_GRADIENT = [2728, 7158, 10775, 23399, 72780]


def satd_pattern(test, log_w):
    """Buffers of tests/satd_tests.c:59-101 (setup_tests)."""
    n = 1 << log_w
    size = n * n
    i = np.arange(size)
    if test == 0:
        return np.zeros(size, np.uint8), np.full(size, 255, np.uint8)
    if test == 1:
        a = 255 * ((((i >> log_w) % 2) + (i % 2)) % 2)
        b = (a + 1) % 2
        return a.astype(np.uint8), b.astype(np.uint8)
    col, row = i % n, i // n
    r = np.array([int(math.sqrt(int(rr) * int(rr) + int(cc) * int(cc))) for rr, cc in zip(row, col)])
    a = 255 // (r + 1)
    return a.astype(np.uint8), (255 - a).astype(np.uint8)


# tests/sad_tests.c:56-76 8x8 pictures (+48 offset, :95-106) and :144-282 expectations for
# uvg_image_calc_sad(pic, ref, 0, 0, X, Y, 8, 8)
SAD_REF_8x8 = np.array([
    1, 2, 2, 2, 2, 2, 2, 3,
    4, 5, 5, 5, 5, 5, 5, 6,
    4, 5, 5, 5, 5, 5, 5, 6,
    4, 5, 5, 5, 5, 5, 5, 6,
    4, 5, 5, 5, 5, 5, 5, 6,
    4, 5, 5, 5, 5, 5, 5, 6,
    4, 5, 5, 5, 5, 5, 5, 6,
    7, 8, 8, 8, 8, 8, 8, 9], np.uint8) + 48
SAD_PIC_8x8 = np.ones(64, np.uint8) + 48
_D = 10  # tests/sad_tests.c:219
SAD_EDGE_CASES = [  # (ref_x, ref_y, expected)
    (-3, -3, 1 * 16 + (2 + 4) * 16 + 5 * 16 - 64),
    (0, -3, (1 + 3) * 4 + 2 * 24 + (4 + 6) * 4 + 5 * 24 - 64),
    (3, -3, 3 * 16 + (2 + 6) * 16 + 5 * 16 - 64),
    (-3, 0, (1 + 7) * 4 + 4 * 24 + (2 + 8) * 4 + 5 * 24 - 64),
    (0, 0, (1 + 3 + 7 + 9) + (2 + 4 + 6 + 8) * 6 + 5 * 36 - 64),
    (3, 0, (3 + 9) * 4 + 6 * 24 + (2 + 8) * 4 + 5 * 24 - 64),
    (-3, 3, 7 * 16 + (4 + 8) * 16 + 5 * 16 - 64),
    (0, 3, (7 + 9) * 4 + 8 * 24 + (4 + 6) * 4 + 5 * 24 - 64),
    (3, 3, 9 * 16 + (6 + 8) * 16 + 5 * 16 - 64),
    (-_D, -_D, 1 * 64 - 64),
    (0, -_D, (1 + 3) * 8 + 2 * 48 - 64),
    (_D, -_D, 3 * 64 - 64),
    (-_D, 0, (1 + 7) * 8 + 4 * 48 - 64),
    (_D, 0, (3 + 9) * 8 + 6 * 48 - 64),
    (-_D, _D, 7 * 64 - 64),
    (0, _D, (7 + 9) * 8 + 8 * 48 - 64),
    (_D, _D, 9 * 64 - 64),
]

# tests/sad_tests.c:388-399 - reg_sad shapes tested against a trivial SAD
REG_SAD_DIMS = [(64, 64), (32, 32), (16, 16), (8, 8),
                (64, 32), (32, 64), (32, 16), (16, 32), (16, 8), (8, 16), (8, 4), (4, 8),
                (48, 16), (16, 48), (24, 16), (16, 24), (12, 4), (4, 12)]


def sad_big_pictures():
    """tests/sad_tests.c:108-124: two 64x64 pseudo-random pictures."""
    i = np.arange(64 * 64, dtype=np.int64)
    pic = ((i * i // 32 + i) % 255).astype(np.uint8)
    ref = ((i * i // 16 + i) % 255).astype(np.uint8)
    return pic, ref


def dct_gradient():
    """tests/dct_tests.c:68 init_gradient-style 64x64 radial gradient, as int16."""
    y, x = np.mgrid[0:64, 0:64]
    r = np.sqrt(x * x + y * y).astype(np.int64)
    return (255 // (r + 1)).astype(np.int16)


def intra_sad_gradient(log_w):
    """tests/intra_sad_tests.c:60-70,104-110: init_gradient(3, 1, width, 1) vs flat 128."""
    n = 1 << log_w
    y, x = np.mgrid[0:n, 0:n]
    val = (np.sqrt((3 - x) ** 2 + (1 - y) ** 2) + 0.5 + 1).astype(np.int64)
    return np.clip(val, 0, 255).astype(np.uint8).ravel(), np.full(n * n, 128, np.uint8)


def coeff_sum_case():
    """tests/coeff_sum_tests.c:44-56: 4096 coeffs INT16_MIN, +16, ... and the closed form."""
    data = (np.arange(4096, dtype=np.int64) * 16 - 32768).astype(np.int16)
    expected = 2048 * (16 + 32768) // 2 + 2048 * 2047 * 16 // 2
    return data, expected
