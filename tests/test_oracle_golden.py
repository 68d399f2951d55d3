"""Oracle (oracle/uvg_oracle.c) against the reference's own known-answer tests.

CPU only. These are the vectors SURVEY.md 8(c) lists: satd golden constants, the 17
frame-edge SAD answers, black/white + gradient SADs, coeff_abs_sum closed form.
"""
import ctypes as C
import numpy as np
import pytest

from _libs import load_oracle, ptr
from golden import reference_kats as K

O = load_oracle(8)


@pytest.mark.parametrize("log_w", [2, 3, 4, 5, 6])
@pytest.mark.parametrize("test,expected", [(0, K.SATD_BLACK_WHITE), (1, K.SATD_CHECKERS),
                                           (2, K.SATD_GRADIENT)])
def test_satd_golden(log_w, test, expected):
    a, b = K.satd_pattern(test, log_w)
    n = 1 << log_w
    r1 = O.orc_satd_nxn(n, ptr(a), ptr(b))
    r2 = O.orc_satd_nxn(n, ptr(b), ptr(a))
    assert r1 == r2 == expected[log_w - 2]


@pytest.mark.parametrize("rx,ry,expected", K.SAD_EDGE_CASES)
def test_image_sad_edges(rx, ry, expected):
    got = O.orc_image_calc_sad(ptr(K.SAD_PIC_8x8), ptr(K.SAD_REF_8x8), 8, 8, 8, 0, 0, rx, ry, 8, 8)
    assert got == expected


@pytest.mark.parametrize("w,h", K.REG_SAD_DIMS)
def test_reg_sad_shapes(w, h):
    pic, ref = K.sad_big_pictures()
    exp = int(np.abs(pic.reshape(64, 64)[:h, :w].astype(int) - ref.reshape(64, 64)[:h, :w]).sum())
    assert O.orc_reg_sad(ptr(pic), ptr(ref), w, h, 64, 64) == exp
    z = np.zeros(64 * 64, np.uint8); m = np.full(64 * 64, 255, np.uint8)
    assert O.orc_reg_sad(ptr(z), ptr(m), w, h, 64, 64) == 255 * w * h  # overflow case


@pytest.mark.parametrize("log_w", [2, 3, 4, 5, 6])
def test_intra_sad(log_w):
    n = 1 << log_w
    z = np.zeros(n * n, np.uint8); m = np.full(n * n, 255, np.uint8)
    assert O.orc_sad_nxn(n, ptr(z), ptr(m)) == 255 * n * n == O.orc_sad_nxn(n, ptr(m), ptr(z))
    a, b = K.intra_sad_gradient(log_w)
    exp = int(np.abs(a.astype(int) - b).sum())
    assert O.orc_sad_nxn(n, ptr(a), ptr(b)) == exp == O.orc_sad_nxn(n, ptr(b), ptr(a))


def test_coeff_abs_sum():
    data, expected = K.coeff_sum_case()
    assert O.orc_coeff_abs_sum(ptr(data), data.size) == expected


def test_transform_roundtrip_property():
    """idct(dct(x)) ~= x for DCT-2 (within rounding) - sanity of matrices and shifts."""
    rng = np.random.default_rng(7)
    for n in (4, 8, 16, 32):
        x = rng.integers(-255, 256, n * n).astype(np.int16)
        c = np.zeros(n * n, np.int16); y = np.zeros(n * n, np.int16)
        O.orc_dct_nxn(n, 8, ptr(x), ptr(c))
        O.orc_idct_nxn(n, 8, ptr(c), ptr(y))
        assert np.abs(x.astype(int) - y).max() <= 6


def test_matrices_are_near_orthogonal():
    for t in (0, 1, 2):
        for n in (4, 8, 16, 32):
            T = np.zeros(n * n, np.int16)
            assert O.orc_tr_matrix(t, n, ptr(T))
            T = T.reshape(n, n).astype(np.int64)
            G = T @ T.T
            d = np.diag(G)
            assert np.all(np.abs(d - 64 * 64 * n) < 0.01 * 64 * 64 * n)
            off = G - np.diag(d)
            assert np.abs(off).max() < 0.02 * 64 * 64 * n
