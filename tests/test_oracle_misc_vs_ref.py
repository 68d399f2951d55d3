"""Pins orc_crc32c_nxn and orc_bipred_average (the non-distortion leftovers of the picture strategy table)
to the unmodified reference (picture-generic.c:1133-1247, :1371-1443). CPU only."""
import ctypes as C

import numpy as np
import pytest

from _libs import load_oracle, load_ref, ref_available, ptr, pix_dtype

pytestmark = pytest.mark.skipif(not ref_available(8), reason="oracle/_ref not built")
vp, ci = C.c_void_p, C.c_int


@pytest.mark.parametrize("bd", [8, 10])
def test_crc32c_and_bipred(bd):
    O, R = load_oracle(bd), load_ref(bd)
    O.orc_crc32c_nxn.argtypes = [ci, vp, C.c_uint32]; O.orc_crc32c_nxn.restype = C.c_uint32
    R.H.refh_crc32c.argtypes = [ci, vp, C.c_uint32]; R.H.refh_crc32c.restype = C.c_uint32
    O.orc_bipred_average.argtypes = [vp, ci, vp, ci, vp, ci, ci, ci]
    R.H.refh_bipred_average.argtypes = [vp, vp, ci, vp, ci, ci, ci]
    rng = np.random.default_rng(6600 + bd)
    hi = (1 << bd) - 1
    plane = rng.integers(0, hi + 1, (40, 48)).astype(pix_dtype(bd))
    for n in (4, 8):
        for _ in range(50):
            y, x = int(rng.integers(0, 40 - n)), int(rng.integers(0, 48 - n))
            p = plane.ctypes.data + (y * 48 + x) * plane.itemsize
            assert O.orc_crc32c_nxn(n, p, 48) == R.H.refh_crc32c(n, p, 48)
    for (w, h) in ((4, 4), (8, 8), (16, 8), (8, 16), (32, 32), (64, 64), (4, 16)):
        for (i0, i1) in ((0, 0), (1, 1), (0, 1), (1, 0)):
            srcs = []
            for is_im in (i0, i1):
                if is_im:   # 14-bit intermediates incl. values that round outside the pixel range
                    srcs.append(rng.integers(-2000, (hi << (14 - bd)) + 2000, w * h).astype(np.int16))
                else:
                    srcs.append(rng.integers(0, hi + 1, w * h).astype(pix_dtype(bd)))
            d1 = np.zeros((h, 64), pix_dtype(bd)); d2 = np.zeros((h, 64), pix_dtype(bd))
            O.orc_bipred_average(ptr(d1), 64, ptr(srcs[0]), i0, ptr(srcs[1]), i1, w, h)
            R.H.refh_bipred_average(ptr(d2), ptr(srcs[0]), i0, ptr(srcs[1]), i1, w, h)
            assert (d1[:, :w] == d2[:, :w]).all(), (bd, w, h, i0, i1)


@pytest.mark.parametrize("bd", [8, 10])
def test_pixel_var(bd):
    """pixel_var (picture-generic.c:1334-1357): bit-exact doubles, incl. a picture-sized array."""
    O, R = load_oracle(bd), load_ref(bd)
    O.orc_pixel_var.argtypes = [vp, C.c_uint32]; O.orc_pixel_var.restype = C.c_double
    f = R.fn("pixel_var", C.c_double, [vp, C.c_uint32])
    rng = np.random.default_rng(6700 + bd)
    for n in (1, 7, 64, 4096, 1024 * 1024 + 3):
        a = rng.integers(0, 1 << bd, n).astype(pix_dtype(bd))
        assert O.orc_pixel_var(ptr(a), n) == f(ptr(a), n)
        flat = np.full(n, (1 << bd) - 1, pix_dtype(bd))
        assert O.orc_pixel_var(ptr(flat), n) == f(ptr(flat), n) == 0.0


def test_frame_pass_with_the_references_avx2_strategies_equals_generic():
    """bench.py reports the CPU frame pass with the reference's AVX2 strategies as context next to the generic-C
    baseline: same outputs on the same picture rows (the harness only repoints the strategy table)."""
    import os, sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    try:
        arm = bench.CpuArm(60)
    except Exception as exc:
        pytest.skip("oracle/_ref not built: %s" % exc)
    arm.frame_setup(1)
    arm.frame_run(2)
    keep = {k: [x.copy() for x in v] for k, v in arm.f_arrays.items()}
    try:
        if arm.frame_strategy("avx2") <= 0:
            pytest.skip("no AVX2 strategies registered on this CPU")
        arm.frame_run(2)
        for k, v in arm.f_arrays.items():
            for a, b in zip(v, keep[k]):
                assert (a == b).all(), k
    finally:
        arm.frame_strategy("generic")
