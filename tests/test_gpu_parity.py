"""GPU parity: the CUDA path (through the C ABI of libuvg266_b200.so) against the oracle
on the same seeded inputs, the reference's known-answer tests, and - when oracle/_ref is
present - the unmodified reference itself. Bit-exact everywhere (integer domain).

Run on the B200 box: python -m pytest tests -m gpu
"""
import ctypes as C

import numpy as np
import pytest

import uvg266_b200 as U
from _libs import load_oracle, load_ref, ptr, pix_dtype, OrcQuantParams
from golden import reference_kats as K

pytestmark = pytest.mark.gpu
BDS = [8, 10]
vp = C.c_void_p


@pytest.fixture(scope="module", autouse=True)
def _init():
    U.init(0)
    U.lib().uvg_b200_set_dropin_bitdepth(10)


def rnd_pix(rng, bd, n, extreme=False):
    hi = 1 << bd
    if extreme:
        return rng.choice(np.array([0, hi - 1]), n).astype(pix_dtype(bd))
    return rng.integers(0, hi, n).astype(pix_dtype(bd))


# --------------------------------------------------------------------------------------
# reference known-answer tests, through the drop-in functions
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("log_w", [2, 3, 4, 5, 6])
@pytest.mark.parametrize("test,expected", [(0, K.SATD_BLACK_WHITE), (1, K.SATD_CHECKERS), (2, K.SATD_GRADIENT)])
def test_kat_satd(log_w, test, expected):
    a, b = K.satd_pattern(test, log_w)
    n = 1 << log_w
    f = U.dropin("satd_%dx%d" % (n, n), 8)
    assert f(ptr(a), ptr(b)) == f(ptr(b), ptr(a)) == expected[log_w - 2]


@pytest.mark.parametrize("log_w", [2, 3, 4, 5, 6])
def test_kat_intra_sad(log_w):
    n = 1 << log_w
    f = U.dropin("sad_%dx%d" % (n, n), 8)
    z = np.zeros(n * n, np.uint8); m = np.full(n * n, 255, np.uint8)
    assert f(ptr(z), ptr(m)) == 255 * n * n == f(ptr(m), ptr(z))
    a, b = K.intra_sad_gradient(log_w)
    assert f(ptr(a), ptr(b)) == int(np.abs(a.astype(int) - b).sum()) == f(ptr(b), ptr(a))


@pytest.mark.parametrize("w,h", K.REG_SAD_DIMS)
def test_kat_reg_sad_shapes(w, h):
    f = U.dropin("reg_sad", 8)
    pic, ref = K.sad_big_pictures()
    exp = int(np.abs(pic.reshape(64, 64)[:h, :w].astype(int) - ref.reshape(64, 64)[:h, :w]).sum())
    assert f(ptr(pic), ptr(ref), w, h, 64, 64) == exp
    z = np.zeros(64 * 64, np.uint8); m = np.full(64 * 64, 255, np.uint8)
    assert f(ptr(z), ptr(m), w, h, 64, 64) == 255 * w * h
    g = U.lib().uvg_b200_get_optimized_sad_8(w)
    if w in (4, 8, 12, 16, 24, 32, 64):
        fn = C.CFUNCTYPE(C.c_uint32, vp, vp, C.c_int32, C.c_uint32, C.c_uint32)(g)
        assert fn(ptr(pic), ptr(ref), h, 64, 64) == exp
    else:
        assert not g


def test_empty_blocks_cost_nothing():
    """image_interpolated_sad (image.c:404-416) calls reg_sad / the width-specialised SADs / ver_sad with a
    height of 0 when the motion vector points entirely above or below the frame: the reference's loops do
    not run and return 0 (found by the --me=tz identity case)."""
    pic, ref = K.sad_big_pictures()
    L = U.lib()
    assert U.dropin("reg_sad", 8)(ptr(pic), ptr(ref), 16, 0, 64, 64) == 0
    assert U.dropin("reg_sad", 8)(ptr(pic), ptr(ref), 0, 16, 64, 64) == 0
    fn = C.CFUNCTYPE(C.c_uint32, vp, vp, C.c_int32, C.c_uint32, C.c_uint32)(L.uvg_b200_get_optimized_sad_8(16))
    assert fn(ptr(pic), ptr(ref), 0, 64, 64) == 0
    assert U.dropin("ver_sad", 8)(ptr(pic), ptr(ref), 16, 0, 64) == 0
    assert U.dropin("hor_sad", 8)(ptr(pic), ptr(ref), 16, 0, 64, 64, 4, 0) == 0


@pytest.mark.parametrize("bd", BDS)
def test_pixel_var_bit_exact_doubles(bd):
    """pixel_var: the variance the reference computes in double with index-order accumulation, reproduced bit for
    bit (drop-in on arrays up to picture size, batched entry point on 64x64 CTUs)."""
    import torch
    O = load_oracle(bd)
    O.orc_pixel_var.argtypes = [vp, C.c_uint32]; O.orc_pixel_var.restype = C.c_double
    L = U.lib()
    f = getattr(L, "uvg_b200_pixel_var_%s" % U.suffix(bd))
    rng = np.random.default_rng(1950 + bd)
    for n in (1, 5, 64, 4095, 4096, 4097, 64 * 64, 832 * 480, 1920 * 1080 + 1):
        a = rnd_pix(rng, bd, n, extreme=(n == 4097))
        assert f(ptr(a), n) == O.orc_pixel_var(ptr(a), n), n
    ctus = rnd_pix(rng, bd, 300 * 4096)
    d = torch.from_numpy(ctus.view(np.uint8) if bd == 8 else ctus.view(np.int16)).cuda()
    out = torch.zeros(300, dtype=torch.float64, device="cuda")
    U.check(L.uvg_b200_pixel_var_dev(bd, d.data_ptr(), 4096, 4096, out.data_ptr(), 300, None))
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    for i in range(300):
        blk = np.ascontiguousarray(ctus[i * 4096:(i + 1) * 4096])
        assert got[i] == O.orc_pixel_var(ptr(blk), 4096), i


def test_kat_coeff_abs_sum():
    data, expected = K.coeff_sum_case()
    assert U.lib().uvg_b200_coeff_abs_sum(ptr(data), data.size) == expected


def test_kat_image_sad_edges_via_search():
    """tests/sad_tests.c:144-282 through the batched MV-search kernel (edge replication)."""
    import torch
    L = U.lib()
    pic = torch.from_numpy(K.SAD_PIC_8x8.copy()).cuda(); ref = torch.from_numpy(K.SAD_REF_8x8.copy()).cuda()
    r = 10
    side = 2 * r + 1
    zero = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = torch.zeros(side * side, dtype=torch.int32, device="cuda")
    U.check(L.uvg_b200_sad_search_dev(8, 8, 8, pic.data_ptr(), ref.data_ptr(), 8, 8, 8, zero.data_ptr(),
                                      zero.data_ptr(), zero.data_ptr(), zero.data_ptr(), r, out.data_ptr(), 1, None))
    torch.cuda.synchronize()
    o = out.cpu().numpy().reshape(side, side)
    for rx, ry, expected in K.SAD_EDGE_CASES:
        assert o[ry + r, rx + r] == expected, (rx, ry)


# --------------------------------------------------------------------------------------
# drop-in functions vs oracle on random inputs
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("bd", BDS)
def test_dropin_picture_nxn(bd):
    O = load_oracle(bd)
    rng = np.random.default_rng(1000 + bd)
    for n in (4, 8, 16, 32, 64):
        sad, satd = U.dropin("sad_%dx%d" % (n, n), bd), U.dropin("satd_%dx%d" % (n, n), bd)
        for it in range(6):
            a = rnd_pix(rng, bd, n * n, extreme=(it == 0)); b = rnd_pix(rng, bd, n * n, extreme=(it == 0))
            assert sad(ptr(a), ptr(b)) == O.orc_sad_nxn(n, ptr(a), ptr(b))
            assert satd(ptr(a), ptr(b)) == O.orc_satd_nxn(n, ptr(a), ptr(b))
            if True:
                preds = rnd_pix(rng, bd, 1024 + max(1024, n * n))    # 64x64 dual: overlapping 4096-pixel reads
                c1 = np.zeros(2, np.uint32); c2 = np.zeros(2, np.uint32)
                U.dropin("sad_%dx%d_dual" % (n, n), bd)(ptr(preds), ptr(a), 2, ptr(c1))
                O.orc_sad_nxn_dual(n, ptr(preds), ptr(a), ptr(c2)); assert (c1 == c2).all()
                U.dropin("satd_%dx%d_dual" % (n, n), bd)(ptr(preds), ptr(a), 2, ptr(c1))
                O.orc_satd_nxn_dual(n, ptr(preds), ptr(a), ptr(c2)); assert (c1 == c2).all()


@pytest.mark.parametrize("bd", BDS)
def test_dropin_picture_strided(bd):
    O = load_oracle(bd)
    rng = np.random.default_rng(1100 + bd)
    S1, S2 = 80, 72
    dims = [(w, h) for w in (4, 8, 12, 16, 24, 32, 48, 64) for h in (4, 8, 12, 16, 32, 64)]
    for (w, h) in dims:
        ext = (w == 64 and h == 64)
        a = rnd_pix(rng, bd, S1 * 66, ext); b = rnd_pix(rng, bd, S2 * 66, ext)
        # start at an odd offset: MV-displaced blocks are arbitrarily aligned
        oa, ob = int(rng.integers(0, 8)), int(rng.integers(0, 8))
        pa = a[oa:].ctypes.data_as(vp); pb = b[ob:].ctypes.data_as(vp)
        assert U.dropin("reg_sad", bd)(pa, pb, w, h, S1, S2) == O.orc_reg_sad(pa, pb, w, h, S1, S2)
        assert U.dropin("satd_any_size", bd)(w, h, pa, S1, pb, S2) == O.orc_satd_any_size(w, h, pa, S1, pb, S2)
        assert U.dropin("satd_any_size_vtm", bd)(w, h, pa, S1, pb, S2) == O.orc_satd_any_size_vtm(w, h, pa, S1, pb, S2)
        assert U.dropin("pixels_calc_ssd", bd)(pa, pb, S1, S2, w, h) == O.orc_pixels_calc_ssd(pa, pb, S1, S2, w, h)
        assert U.dropin("ver_sad", bd)(pa, pb, w, h, S1) == O.orc_ver_sad(pa, pb, w, h, S1)
        for (l, r) in ((0, 0), (3, 0), (0, 2), (w - 1, 0), (0, w - 1)):
            assert U.dropin("hor_sad", bd)(pa, pb, w, h, S1, S2, l, r) == O.orc_hor_sad(pa, pb, w, h, S1, S2, l, r)
        r1 = np.zeros(w * h, np.int16); r2 = np.zeros(w * h, np.int16)
        U.dropin("generate_residual", bd)(pa, pb, ptr(r1), w, h, S1, S2)
        O.orc_generate_residual(pa, pb, ptr(r2), w, h, S1, S2)
        assert (r1 == r2).all()
        P = [rnd_pix(rng, bd, S2 * 66) for _ in range(4)]
        arr = (vp * 4)(*[p.ctypes.data for p in P])
        c1 = np.zeros(4, np.uint32); c2 = np.zeros(4, np.uint32)
        U.dropin("satd_any_size_quad", bd)(w, h, arr, S2, pa, S1, 4, ptr(c1), None)
        O.orc_satd_any_size_quad(w, h, arr, S2, pa, S1, ptr(c2))
        assert (c1 == c2).all(), (w, h)
    for (w, h) in ((2, 2), (2, 4), (4, 2), (2, 8), (6, 6), (2, 16)):
        a = rnd_pix(rng, bd, S1 * 20); b = rnd_pix(rng, bd, S2 * 20)
        assert U.dropin("satd_any_size_vtm", bd)(w, h, ptr(a), S1, ptr(b), S2) == \
            O.orc_satd_any_size_vtm(w, h, ptr(a), S1, ptr(b), S2)


@pytest.mark.parametrize("bd", BDS)
def test_dropin_transforms(bd):
    O = load_oracle(bd)
    L = U.lib()
    rng = np.random.default_rng(1200 + bd)
    for n in (4, 8, 16, 32):
        for amp in (255, 1023, 32767):
            for _ in range(3):
                x = rng.integers(-amp, amp + 1, n * n).astype(np.int16)
                o1 = np.zeros(n * n, np.int16); o2 = np.zeros(n * n, np.int16)
                getattr(L, "uvg_b200_dct_%dx%d" % (n, n))(bd, ptr(x), ptr(o1)); O.orc_dct_nxn(n, bd, ptr(x), ptr(o2))
                assert (o1 == o2).all()
                getattr(L, "uvg_b200_idct_%dx%d" % (n, n))(bd, ptr(x), ptr(o1)); O.orc_idct_nxn(n, bd, ptr(x), ptr(o2))
                assert (o1 == o2).all()
    for amp in (255, 32767):
        x = rng.integers(-amp, amp + 1, 16).astype(np.int16)
        o1 = np.zeros(16, np.int16); o2 = np.zeros(16, np.int16)
        L.uvg_b200_fast_forward_dst_4x4(bd, ptr(x), ptr(o1)); O.orc_fast_forward_dst_4x4(bd, ptr(x), ptr(o2)); assert (o1 == o2).all()
        L.uvg_b200_fast_inverse_dst_4x4(bd, ptr(x), ptr(o1)); O.orc_fast_inverse_dst_4x4(bd, ptr(x), ptr(o2)); assert (o1 == o2).all()
    # mts: all shapes x type selectors
    cases = 0
    sizes = (1, 2, 4, 8, 16, 32)
    for w in sizes:
        for h in sizes:
            if w * h < 4 or (w == 2 and h == 2):
                continue
            for mts_type in (0, 1, 2, 3, 4):
                for cu_type in (1, 2):
                    for tr_idx in (0, 2, 3, 4, 5):
                        for lfnst in (0, 1):
                            for color in (0, 1):
                                if (w < 4 or h < 4) and ((color == 0 and mts_type != 0) or lfnst):
                                    continue
                                if rng.random() > 0.12:
                                    continue
                                amp = int(rng.choice([255, 1023, 32767]))
                                x = rng.integers(-amp, amp + 1, w * h).astype(np.int16)
                                o1 = np.full(w * h, 77, np.int16); o2 = np.full(w * h, 77, np.int16)
                                L.uvg_b200_mts_dct(bd, color, cu_type, tr_idx, lfnst, lfnst, 0, w, h, ptr(x), ptr(o1), mts_type)
                                O.orc_mts_dct(bd, color, cu_type, tr_idx, lfnst, lfnst, 0, w, h, ptr(x), ptr(o2), mts_type)
                                assert (o1 == o2).all(), ("fwd", w, h, mts_type, cu_type, tr_idx, lfnst, color)
                                L.uvg_b200_mts_idct(bd, color, cu_type, tr_idx, lfnst, lfnst, 0, w, h, ptr(x), ptr(o1), mts_type)
                                O.orc_mts_idct(bd, color, cu_type, tr_idx, lfnst, lfnst, 0, w, h, ptr(x), ptr(o2), mts_type)
                                assert (o1 == o2).all(), ("inv", w, h, mts_type, cu_type, tr_idx, lfnst, color)
                                cases += 1
    assert cases > 150


QUANT_SHAPES = [(4, 4), (8, 8), (16, 16), (32, 32), (8, 4), (4, 8), (16, 4), (4, 16), (16, 8),
                (8, 16), (32, 8), (8, 32), (32, 16), (16, 32), (2, 8), (8, 2), (2, 16), (32, 4)]
# default VVC chroma QP mapping of the reference (encoder.c:141-183) is obtained from
# oracle/_ref when present; otherwise chroma uses the identity clip of uvg_get_scaled_qp.


def chroma_map(bd):
    R = load_ref(bd)
    if R is None:
        return None, None
    p = R.H.refh_chroma_qp_map()
    arr = (C.c_int8 * 64).from_address(p)
    return p, [arr[i] for i in range(64)]


@pytest.mark.parametrize("bd", BDS)
@pytest.mark.parametrize("signhide", [0, 1])
def test_dropin_quant_dequant(bd, signhide):
    O = load_oracle(bd)
    L = U.lib()
    rng = np.random.default_rng(1300 + bd + signhide)
    qptr, qlist = chroma_map(bd)
    n_cases = 0
    for (w, h) in QUANT_SHAPES:
        for qp in (0, 7, 22, 27, 32, 37, 44, 51):
            for slice_i in (0, 1):
                for color in (0, 1, 2):
                    for ts in (0, 1):
                        if rng.random() > 0.25:
                            continue
                        lfnst = int(rng.integers(0, 3)) if (w >= 4 and h >= 4 and rng.random() < 0.3) else 0
                        p = OrcQuantParams(qp, bd, slice_i, signhide, 0, qptr)
                        cfg = U.QuantCfg.make(qp, bd, slice_i, signhide, 0, qlist)
                        amp = int(rng.choice([40, 600, 5000, 32767]))
                        coef = rng.integers(-amp, amp + 1, w * h).astype(np.int16)
                        if rng.random() < 0.3:
                            coef[rng.random(w * h) < 0.8] = 0
                        q1 = np.zeros(w * h, np.int16); q2 = np.zeros(w * h, np.int16)
                        L.uvg_b200_quant(C.byref(cfg), ptr(coef), ptr(q1), w, h, color, 0, 1, ts, lfnst)
                        O.orc_quant(C.byref(p), ptr(coef), ptr(q2), w, h, color, 1, ts, lfnst)
                        assert (q1 == q2).all(), (w, h, qp, slice_i, color, ts, lfnst)
                        d1 = np.zeros(w * h, np.int16); d2 = np.zeros(w * h, np.int16)
                        L.uvg_b200_dequant(C.byref(cfg), ptr(q1), ptr(d1), w, h, color, 1, ts)
                        O.orc_dequant(C.byref(p), ptr(q2), ptr(d2), w, h, color, 1, ts)
                        assert (d1 == d2).all(), (w, h, qp, slice_i, color, ts)
                        n_cases += 1
    assert n_cases > 200


@pytest.mark.parametrize("bd", BDS)
@pytest.mark.parametrize("mts_val", [0, 3, 4])
def test_dropin_quantize_residual(bd, mts_val):
    O = load_oracle(bd)
    L = U.lib()
    rng = np.random.default_rng(1400 + bd + mts_val)
    qptr, qlist = chroma_map(bd)
    f = getattr(L, "uvg_b200_quantize_residual_%s" % U.suffix(bd))
    n_cases = 0
    for (w, h) in [(4, 4), (8, 8), (16, 16), (32, 32), (16, 8), (8, 32), (4, 16), (32, 4)]:
        for qp in (12, 22, 32, 42):
            for color in (0, 1):
                for tr_idx in ((0, 2, 3, 5) if mts_val == 3 else (0,)):
                    for early in (0, 1):
                        for alias in (0, 1):
                            sh = (n_cases >> 1) & 1
                            p = OrcQuantParams(qp, bd, 1, sh, mts_val, qptr)
                            cfg = U.QuantCfg.make(qp, bd, 1, sh, mts_val, qlist)
                            stride = 64 if color == 0 else 32
                            hi = 1 << bd
                            ref = rng.integers(0, hi, stride * h).astype(pix_dtype(bd))
                            noise = int(rng.choice([2, 12, 60, hi]))
                            pred = np.clip(ref.astype(int) + rng.integers(-noise, noise + 1, stride * h), 0, hi - 1).astype(pix_dtype(bd))
                            pred2 = pred.copy()
                            rec1 = pred if alias else np.full(stride * h, 7, pix_dtype(bd))
                            rec2 = pred2 if alias else np.full(stride * h, 7, pix_dtype(bd))
                            co1 = np.zeros(w * h, np.int16); co2 = np.zeros(w * h, np.int16)
                            hc1 = f(C.byref(cfg), 1, tr_idx, w, h, color, 0, stride, stride, ptr(ref), ptr(pred), ptr(rec1), ptr(co1), early)
                            hc2 = O.orc_quantize_residual(C.byref(p), 1, tr_idx, w, h, color, 0, stride, stride,
                                                          ptr(ref), ptr(pred2), ptr(rec2), ptr(co2), early)
                            assert hc1 == hc2
                            assert (co1 == co2).all(), (w, h, qp, color, tr_idx)
                            assert (rec1 == rec2).all(), (w, h, qp, color, tr_idx, early, alias)
                            n_cases += 1
    assert n_cases >= 256


TS_COMBOS = ((1, 1, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, 2), (0, 2, 1))   # (use_trskip, tr_idx, isp_mode)


@pytest.mark.parametrize("bd", BDS)
@pytest.mark.parametrize("signhide", [0, 1])
def test_dropin_quantize_residual_transform_skip_and_isp(bd, signhide):
    """use_trskip = 1 with / without tr_idx == MTS_SKIP, MTS_SKIP alone, and ISP CUs under explicit MTS
    (oracle pinned on the reference for exactly these combinations: tests/test_oracle_vs_ref.py)."""
    O = load_oracle(bd)
    L = U.lib()
    rng = np.random.default_rng(1420 + bd + signhide)
    qptr, qlist = chroma_map(bd)
    f = getattr(L, "uvg_b200_quantize_residual_%s" % U.suffix(bd))
    n_cases = 0
    for (w, h) in [(4, 4), (8, 8), (16, 16), (32, 32), (16, 4), (4, 16), (8, 32), (32, 8)]:
        for qp in (7, 22, 37):
            for color in (0, 1):
                for (trskip, tr_idx, isp) in TS_COMBOS:
                    early = n_cases & 1
                    slice_i = (n_cases >> 1) & 1
                    p = OrcQuantParams(qp, bd, slice_i, signhide, 3, qptr, isp)
                    cfg = U.QuantCfg.make(qp, bd, slice_i, signhide, 3, qlist, isp)
                    stride = 64 if color == 0 else 32
                    hi = 1 << bd
                    ref = rng.integers(0, hi, stride * h).astype(pix_dtype(bd))
                    noise = int(rng.choice([2, 12, 60, hi]))
                    pred = np.clip(ref.astype(int) + rng.integers(-noise, noise + 1, stride * h), 0, hi - 1).astype(pix_dtype(bd))
                    rec1 = np.full(stride * h, 7, pix_dtype(bd)); rec2 = rec1.copy()
                    co1 = np.zeros(w * h, np.int16); co2 = np.zeros(w * h, np.int16)
                    hc1 = f(C.byref(cfg), 1, tr_idx, w, h, color, trskip, stride, stride, ptr(ref), ptr(pred), ptr(rec1), ptr(co1), early)
                    hc2 = O.orc_quantize_residual(C.byref(p), 1, tr_idx, w, h, color, trskip, stride, stride,
                                                  ptr(ref), ptr(pred), ptr(rec2), ptr(co2), early)
                    assert hc1 == hc2, (w, h, qp, color, trskip, tr_idx, isp)
                    assert (co1 == co2).all(), (w, h, qp, color, trskip, tr_idx, isp)
                    assert (rec1 == rec2).all(), (w, h, qp, color, trskip, tr_idx, isp, early)
                    n_cases += 1
    assert n_cases >= 250


@pytest.mark.parametrize("bd", BDS)
@pytest.mark.parametrize("signhide", [0, 1])
def test_batched_quantize_residual_transform_skip_and_isp(bd, signhide):
    """The same branches through the batched entry point, i.e. through every kernel family that accepts them
    (registers-only 4x4 / 8x8 take use_trskip themselves; the tensor-core kernels hand transform skip to the
    general kernel; ISP changes the transform pair the plans are built for)."""
    O = load_oracle(bd)
    L = U.lib()
    rng = np.random.default_rng(1830 + bd + signhide)
    hi = 1 << bd
    dt = pix_dtype(bd)
    weights = (28 << 0) | (300 << 16) | (700 << 32) | (1100 << 48)
    for (n, nt) in ((4, 150), (8, 67), (16, 21), (32, 5)):
        for (trskip, tr_idx, isp) in TS_COMBOS:
            qp = int(rng.choice([12, 27, 37]))
            ref = rng.integers(0, hi, nt * n * n).astype(dt)
            pred = np.clip(ref.astype(int) + rng.integers(-30, 31, nt * n * n), 0, hi - 1).astype(dt)
            pred[: n * n * 2] = ref[: n * n * 2]
            rec = np.zeros(nt * n * n, dt)
            coeff = np.zeros(nt * n * n, np.int16); has = np.zeros(nt, np.uint8)
            ssd = np.zeros(nt, np.uint32); cost = np.zeros(nt, np.uint32)
            cfg = U.QuantCfg.make(qp, bd, 1, signhide, 3, None, isp)
            R, P, Rc = (U.Blocks.dense(a.ctypes.data, n, n) for a in (ref, pred, rec))
            U.check(L.uvg_b200_quantize_residual_host(C.byref(cfg), 1, tr_idx, n, n, 0, trskip, 0, C.byref(R), ref.nbytes,
                                                      C.byref(P), pred.nbytes, C.byref(Rc), rec.nbytes, ptr(coeff),
                                                      ptr(has), ptr(ssd), ptr(cost), weights, nt))
            p = OrcQuantParams(qp, bd, 1, signhide, 3, None, isp)
            for t in range(nt):
                sl = slice(t * n * n, (t + 1) * n * n)
                r2 = np.zeros(n * n, dt); c2 = np.zeros(n * n, np.int16)
                rr = np.ascontiguousarray(ref[sl]); pp = np.ascontiguousarray(pred[sl])
                h2 = O.orc_quantize_residual(C.byref(p), 1, tr_idx, n, n, 0, trskip, n, n, ptr(rr), ptr(pp), ptr(r2), ptr(c2), 0)
                assert has[t] == h2, (n, trskip, tr_idx, isp, t)
                assert (coeff[sl] == c2).all(), (n, trskip, tr_idx, isp, t)
                assert (rec[sl] == r2).all(), (n, trskip, tr_idx, isp, t)
                assert ssd[t] == O.orc_pixels_calc_ssd(ptr(rr), ptr(r2), n, n, n, n)
                assert cost[t] == O.orc_fast_coeff_cost(ptr(c2), n, n, weights)


@pytest.mark.parametrize("bd", BDS)
def test_quant_cbcr_residual_dropin_and_batched(bd):
    """JCCR (quant_cbcr_residual, quant-generic.c:241-442): the drop-in on strided planes and the batched
    device entry point on dense TUs, all six joint modes, against the oracle."""
    import torch
    O = load_oracle(bd)
    L = U.lib()
    rng = np.random.default_rng(1450 + bd)
    qptr, qlist = chroma_map(bd)
    f = getattr(L, "uvg_b200_quant_cbcr_residual_%s" % U.suffix(bd))
    hi = 1 << bd
    dt = pix_dtype(bd)
    n_cases = 0
    for (w, h) in ((4, 4), (8, 8), (16, 16), (32, 32), (8, 4), (4, 16), (16, 8)):
        for qp in (17, 32):
            for joint in (1, 2, 3):
                for sign in (0, 1):
                    sh = n_cases & 1
                    early = (n_cases >> 1) & 1 if qp == 32 else 0
                    p = OrcQuantParams(qp, bd, 1, sh, 0, qptr)
                    cfg = U.QuantCfg.make(qp, bd, 1, sh, 0, qlist)
                    stride = 32
                    noise = int(rng.choice([2, 12, 60, hi]))
                    pl = []
                    for _ in range(2):
                        ref = rng.integers(0, hi, stride * h).astype(dt)
                        pl.append((ref, np.clip(ref.astype(int) + rng.integers(-noise, noise + 1, stride * h), 0, hi - 1).astype(dt)))
                    (ur, up), (vr, vq) = pl
                    rec = [np.full(stride * h, 7, dt) for _ in range(4)]
                    co1 = np.zeros(w * h, np.int16); co2 = np.zeros(w * h, np.int16)
                    r1 = f(C.byref(cfg), 1, 0, w, h, joint, sign, stride, stride, ptr(ur), ptr(vr), ptr(up), ptr(vq),
                           ptr(rec[0]), ptr(rec[1]), ptr(co1), early)
                    r2 = O.orc_quant_cbcr_residual(C.byref(p), 1, 0, w, h, joint, sign, stride, stride, ptr(ur), ptr(vr),
                                                   ptr(up), ptr(vq), ptr(rec[2]), ptr(rec[3]), ptr(co2), early)
                    assert r1 == r2, (w, h, qp, joint, sign, early)
                    assert (co1 == co2).all(), (w, h, qp, joint, sign)
                    assert (rec[0] == rec[2]).all() and (rec[1] == rec[3]).all(), (w, h, qp, joint, sign, early)
                    n_cases += 1
    # batched: 300 dense 8x8 TUs per mode, one launch each
    dev = torch.device("cuda", 0)
    n, nt = 8, 300
    cfg = U.QuantCfg.make(27, bd, 1, 0, 0, qlist)
    p = OrcQuantParams(27, bd, 1, 0, 0, qptr)
    tdt = torch.uint8 if bd == 8 else torch.int16
    for joint in (1, 2, 3):
        for sign in (0, 1):
            host = []
            for _ in range(2):
                ref = rng.integers(0, hi, nt * n * n).astype(dt)
                pred = np.clip(ref.astype(int) + rng.integers(-20, 21, ref.size), 0, hi - 1).astype(dt)
                pred[: 3 * n * n] = ref[: 3 * n * n]
                host += [ref, pred]
            ur, up, vr, vq = host
            d = [torch.from_numpy(a.view(np.int16) if bd != 8 else a).to(dev) for a in (ur, vr, up, vq)]
            urec = torch.zeros(nt * n * n, dtype=tdt, device=dev); vrec = torch.zeros_like(urec)
            coeff = torch.zeros(nt * n * n, dtype=torch.int16, device=dev)
            has = torch.zeros(nt, dtype=torch.uint8, device=dev)
            B = [U.Blocks.dense(t.data_ptr(), n, n) for t in d + [urec, vrec]]
            U.check(L.uvg_b200_quant_cbcr_residual_dev(C.byref(cfg), 1, 0, n, n, joint, sign, 0, *[C.byref(b) for b in B],
                                                       coeff.data_ptr(), has.data_ptr(), nt, 0))
            torch.cuda.synchronize()
            g_u, g_v = urec.cpu().numpy().view(dt), vrec.cpu().numpy().view(dt)
            g_c, g_h = coeff.cpu().numpy(), has.cpu().numpy()
            for i in range(nt):
                sl = slice(i * n * n, (i + 1) * n * n)
                ru = np.zeros(n * n, dt); rv = np.zeros(n * n, dt); co = np.zeros(n * n, np.int16)
                r = O.orc_quant_cbcr_residual(C.byref(p), 1, 0, n, n, joint, sign, n, n, ptr(ur[sl].copy()), ptr(vr[sl].copy()),
                                              ptr(up[sl].copy()), ptr(vq[sl].copy()), ptr(ru), ptr(rv), ptr(co), 0)
                assert (1 if r else 0) == int(g_h[i]), (joint, sign, i)
                assert (co == g_c[sl]).all() and (ru == g_u[sl]).all() and (rv == g_v[sl]).all(), (joint, sign, i)
    assert n_cases >= 80


def test_dropin_coeff_costs():
    O = load_oracle(8)
    L = U.lib()
    rng = np.random.default_rng(1500)
    for n in (4, 8, 16, 32):
        for _ in range(5):
            c = rng.integers(-6, 7, n * n).astype(np.int16)
            wts = int(rng.integers(0, 1 << 62))
            assert L.uvg_b200_fast_coeff_cost(ptr(c), n, n, wts) == O.orc_fast_coeff_cost(ptr(c), n, n, wts)
            big = rng.integers(-32768, 32768, n * n).astype(np.int16)
            assert L.uvg_b200_coeff_abs_sum(ptr(big), big.size) == O.orc_coeff_abs_sum(ptr(big), big.size)


# --------------------------------------------------------------------------------------
# batched entry points
# --------------------------------------------------------------------------------------
@pytest.mark.parametrize("bd", BDS)
def test_batched_distortion_host(bd):
    """k candidates per block, dense and picture-tiled layouts, every metric."""
    O = load_oracle(bd)
    L = U.lib()
    rng = np.random.default_rng(1600 + bd)
    dt = pix_dtype(bd)
    isz = np.dtype(dt).itemsize
    W, H = 128, 64
    plane = rnd_pix(rng, bd, W * H)
    for (w, h, k) in ((8, 8, 5), (16, 16, 3), (4, 4, 7), (32, 32, 2), (64, 64, 1), (16, 8, 4), (8, 32, 2), (12, 16, 2)):
        if W % w:
            nb = 6
            a_desc = U.Blocks.dense(plane.ctypes.data, w, h)
        else:
            nb = (W // w) * (H // h)
            a_desc = U.Blocks.plane(plane.ctypes.data, W, W, w, h)
        cands = rnd_pix(rng, bd, nb * k * w * h)
        b_desc = U.Blocks.dense(cands.ctypes.data, w, h)
        for metric in (U.SAD, U.SAD_NXN, U.SATD, U.SSD, U.SATD_VTM, U.SATD_QUAD):
            out = np.zeros(nb * k, np.uint32)
            U.check(L.uvg_b200_distortion_host(metric, bd, w, h, C.byref(a_desc), plane.nbytes, C.byref(b_desc),
                                               cands.nbytes, ptr(out), nb, k))
            for i in range(nb):
                if a_desc.blocks_per_row == 1:
                    ao, sa = i * w * h, w
                else:
                    ao, sa = (i // a_desc.blocks_per_row) * h * W + (i % a_desc.blocks_per_row) * w, W
                pa = plane[ao:].ctypes.data_as(vp)
                for c in range(k):
                    pb = cands[(i * k + c) * w * h:].ctypes.data_as(vp)
                    if metric == U.SAD:
                        e = O.orc_reg_sad(pa, pb, w, h, sa, w)
                    elif metric == U.SAD_NXN:
                        e = O.orc_reg_sad(pa, pb, w, h, sa, w) >> (bd - 8)
                    elif metric == U.SATD:
                        e = O.orc_satd_any_size(w, h, pa, sa, pb, w)
                    elif metric == U.SSD:
                        e = O.orc_pixels_calc_ssd(pa, pb, sa, w, w, h)
                    elif metric == U.SATD_VTM:
                        e = O.orc_satd_any_size_vtm(w, h, pa, sa, pb, w)
                    else:
                        P = (vp * 4)(*([pb.value] * 4))
                        q = np.zeros(4, np.uint32)
                        O.orc_satd_any_size_quad(w, h, P, w, pa, sa, ptr(q))
                        e = int(q[0])
                    assert out[i * k + c] == e, (metric, w, h, i, c)


@pytest.mark.parametrize("bd", BDS)
def test_batched_rough_cost(bd):
    O = load_oracle(bd)
    L = U.lib()
    rng = np.random.default_rng(1700 + bd)
    W, H = 128, 64
    plane = rnd_pix(rng, bd, W * H)
    for (n, k) in ((4, 6), (8, 23), (16, 24), (32, 5), (64, 2), (8, 67), (16, 31), (32, 16), (8, 1), (8, 15)):
        nb = (W // n) * (H // n)
        a_desc = U.Blocks.plane(plane.ctypes.data, W, W, n, n)
        # predictions close to the source, like real intra candidates
        cands = np.zeros(nb * k * n * n, pix_dtype(bd))
        for i in range(nb):
            y0, x0 = (i // (W // n)) * n, (i % (W // n)) * n
            blk = plane.reshape(H, W)[y0:y0 + n, x0:x0 + n].astype(int)
            for c in range(k):
                noise = rng.integers(-20 * (c + 1), 20 * (c + 1) + 1, (n, n))
                cands[(i * k + c) * n * n:(i * k + c + 1) * n * n] = np.clip(blk + noise, 0, (1 << bd) - 1).ravel()
        b_desc = U.Blocks.dense(cands.ctypes.data, n, n)
        out = np.zeros(nb * k, np.uint32)
        U.check(L.uvg_b200_rough_cost_host(bd, n, C.byref(a_desc), plane.nbytes, C.byref(b_desc), cands.nbytes,
                                           ptr(out), nb, k))
        for i in range(0, nb, max(1, nb // 16)):
            y0, x0 = (i // (W // n)) * n, (i % (W // n)) * n
            orig = np.ascontiguousarray(plane.reshape(H, W)[y0:y0 + n, x0:x0 + n]).ravel()
            for c in range(k):
                pred = cands[(i * k + c) * n * n:(i * k + c + 1) * n * n]
                if n <= 32:
                    pb = np.zeros(2048, pix_dtype(bd)); pb[:n * n] = pred; pb[1024:1024 + n * n] = pred
                    s = np.zeros(2, np.uint32); t = np.zeros(2, np.uint32)
                    O.orc_satd_nxn_dual(n, ptr(pb), ptr(orig), ptr(s)); O.orc_sad_nxn_dual(n, ptr(pb), ptr(orig), ptr(t))
                    e = min(int(s[0]), 2 * int(t[0]))
                else:
                    e = min(O.orc_satd_nxn(n, ptr(orig), ptr(pred)), 2 * O.orc_sad_nxn(n, ptr(orig), ptr(pred)))
                assert out[i * k + c] == e, (n, i, c)


@pytest.mark.parametrize("bd", BDS)
@pytest.mark.parametrize("signhide", [0, 1])
def test_batched_quantize_residual_and_transform(bd, signhide):
    """signhide=0 takes the shape-specialised kernels (registers-only 4x4/8x8, tensor-core
    16x16/32x32), signhide=1 the general shared-memory kernel."""
    O = load_oracle(bd)
    L = U.lib()
    rng = np.random.default_rng(1800 + bd + signhide)
    hi = 1 << bd
    dt = pix_dtype(bd)
    weights = (28 << 0) | (300 << 16) | (700 << 32) | (1100 << 48)
    for (n, nt) in ((4, 300), (8, 131), (16, 40), (32, 9)):
        for qp in (22, 37):
            ref = rng.integers(0, hi, nt * n * n).astype(dt)
            pred = np.clip(ref.astype(int) + rng.integers(-30, 31, nt * n * n), 0, hi - 1).astype(dt)
            pred[: n * n * 3] = ref[: n * n * 3]           # a few zero-residual TUs: has_coeffs == 0
            rec = np.zeros(nt * n * n, dt)
            coeff = np.zeros(nt * n * n, np.int16); has = np.zeros(nt, np.uint8)
            ssd = np.zeros(nt, np.uint32); cost = np.zeros(nt, np.uint32)
            cfg = U.QuantCfg.make(qp, bd, 1, signhide, 0)
            R, P, Rc = (U.Blocks.dense(a.ctypes.data, n, n) for a in (ref, pred, rec))
            U.check(L.uvg_b200_quantize_residual_host(C.byref(cfg), 1, 0, n, n, 0, 0, 0, C.byref(R), ref.nbytes,
                                                      C.byref(P), pred.nbytes, C.byref(Rc), rec.nbytes, ptr(coeff),
                                                      ptr(has), ptr(ssd), ptr(cost), weights, nt))
            p = OrcQuantParams(qp, bd, 1, signhide, 0, None)
            for t in range(nt):
                sl = slice(t * n * n, (t + 1) * n * n)
                r2 = np.zeros(n * n, dt); c2 = np.zeros(n * n, np.int16)
                rr = np.ascontiguousarray(ref[sl]); pp = np.ascontiguousarray(pred[sl])
                h2 = O.orc_quantize_residual(C.byref(p), 1, 0, n, n, 0, 0, n, n, ptr(rr), ptr(pp), ptr(r2), ptr(c2), 0)
                assert has[t] == h2
                assert (coeff[sl] == c2).all(), (n, qp, t)
                assert (rec[sl] == r2).all(), (n, qp, t)
                assert ssd[t] == O.orc_pixels_calc_ssd(ptr(rr), ptr(r2), n, n, n, n)
                assert cost[t] == O.orc_fast_coeff_cost(ptr(c2), n, n, weights)
            assert has[:3].sum() == 0
        # batched transforms, both directions
        x = rng.integers(-1023, 1024, nt * n * n).astype(np.int16)
        y = np.zeros_like(x); z = np.zeros_like(x)
        for th, tv in ((0, 0), (2, 2), (1, 2)):
            sk = 16 if (n == 32 and th != 0) else 0
            skv = 16 if (n == 32 and tv != 0) else 0
            U.check(L.uvg_b200_transform_host(0, bd, n, n, th, tv, sk, skv, ptr(x), ptr(y), nt))
            U.check(L.uvg_b200_transform_host(1, bd, n, n, th, tv, sk, skv, ptr(y), ptr(z), nt))
            tr_idx = {(0, 0): 0, (2, 2): 2, (1, 2): 3}[(th, tv)]
            for t in range(0, nt, max(1, nt // 8)):
                sl = slice(t * n * n, (t + 1) * n * n)
                xx = np.ascontiguousarray(x[sl]); o = np.zeros(n * n, np.int16); o2 = np.zeros(n * n, np.int16)
                O.orc_mts_dct(bd, 0, 1, tr_idx, 0, 0, 0, n, n, ptr(xx), ptr(o), 3)
                assert (y[sl] == o).all(), (n, th, tv)
                O.orc_mts_idct(bd, 0, 1, tr_idx, 0, 0, 0, n, n, ptr(o), ptr(o2), 3)
                assert (z[sl] == o2).all(), (n, th, tv)


@pytest.mark.parametrize("bd", BDS)
def test_sad_search_vs_oracle(bd):
    import torch
    O = load_oracle(bd)
    L = U.lib()
    rng = np.random.default_rng(1900 + bd)
    W, H = 96, 64
    cur = rnd_pix(rng, bd, W * H); ref = rnd_pix(rng, bd, W * H)
    tdt = torch.uint8 if bd == 8 else torch.int16
    d_cur = torch.from_numpy(cur.view(np.uint8 if bd == 8 else np.int16)).cuda()
    d_ref = torch.from_numpy(ref.view(np.uint8 if bd == 8 else np.int16)).cuda()
    for (w, h, r) in ((8, 8, 4), (16, 16, 8), (32, 16, 3), (64, 64, 2)):
        nb = 5
        cx = rng.integers(0, W - w + 1, nb).astype(np.int32); cy = rng.integers(0, H - h + 1, nb).astype(np.int32)
        mx = rng.integers(-40, 41, nb).astype(np.int32); my = rng.integers(-40, 41, nb).astype(np.int32)
        side = 2 * r + 1
        out = torch.zeros(nb * side * side, dtype=torch.int32, device="cuda")
        t = [torch.from_numpy(a).cuda() for a in (cx, cy, mx, my)]
        U.check(L.uvg_b200_sad_search_dev(bd, w, h, d_cur.data_ptr(), d_ref.data_ptr(), W, H, W, t[0].data_ptr(),
                                          t[1].data_ptr(), t[2].data_ptr(), t[3].data_ptr(), r, out.data_ptr(), nb, None))
        torch.cuda.synchronize()
        o = out.cpu().numpy().reshape(nb, side, side)
        for j in range(nb):
            for dy in range(-r, r + 1):
                for dx in range(-r, r + 1):
                    e = O.orc_image_calc_sad(ptr(cur), ptr(ref), W, H, W, int(cx[j]), int(cy[j]),
                                             int(cx[j] + mx[j] + dx), int(cy[j] + my[j] + dy), w, h)
                    assert o[j, dy + r, dx + r] == e, (w, h, j, dx, dy)


def test_full_size_properties():
    """BASELINE-size batches: size-independent properties instead of a CPU replay.
    2^15 pairs of 64x64: SAD(a,a)=0, SAD symmetric, SAD(a,0)=sum(a); SATD(a,a)=0; linearity of
    SSD under swapping; and a checksum of the batch against numpy for SAD."""
    import torch
    L = U.lib()
    n, N = 64, 1 << 15
    g = torch.Generator(device="cuda"); g.manual_seed(1234)
    a = torch.randint(0, 256, (N, n * n), dtype=torch.uint8, device="cuda", generator=g)
    b = torch.randint(0, 256, (N, n * n), dtype=torch.uint8, device="cuda", generator=g)
    zero = torch.zeros_like(a)
    out1 = torch.zeros(N, dtype=torch.int32, device="cuda"); out2 = torch.zeros_like(out1)
    A, B, Z = (U.Blocks.dense(t.data_ptr(), n, n) for t in (a, b, zero))
    for metric in (U.SAD, U.SATD, U.SSD):
        U.check(L.uvg_b200_distortion_dev(metric, 8, n, n, C.byref(A), C.byref(A), out1.data_ptr(), N, 1, None))
        assert int(out1.abs().max()) == 0
        U.check(L.uvg_b200_distortion_dev(metric, 8, n, n, C.byref(A), C.byref(B), out1.data_ptr(), N, 1, None))
        U.check(L.uvg_b200_distortion_dev(metric, 8, n, n, C.byref(B), C.byref(A), out2.data_ptr(), N, 1, None))
        assert torch.equal(out1, out2)
    U.check(L.uvg_b200_distortion_dev(U.SAD, 8, n, n, C.byref(A), C.byref(Z), out1.data_ptr(), N, 1, None))
    assert torch.equal(out1.long(), a.long().sum(1))
    U.check(L.uvg_b200_distortion_dev(U.SAD, 8, n, n, C.byref(A), C.byref(B), out1.data_ptr(), N, 1, None))
    assert torch.equal(out1.long(), (a.int() - b.int()).abs().sum(1).long())
    U.check(L.uvg_b200_distortion_dev(U.SSD, 8, n, n, C.byref(A), C.byref(B), out1.data_ptr(), N, 1, None))
    d = (a.int() - b.int()).long()
    assert torch.equal(out1.long(), (d * d).sum(1))


@pytest.mark.parametrize("mode", ["staged", "mma", "umma", "tma", "tc5"])
def test_rough_cost_kernel_variants(mode):
    """The interchangeable rough-cost kernels (packed ALU, mma.sync IMMA, tcgen05/TMEM, TMA-fed) are
    selected once per process by UVG266_B200_ROUGH: run the batched parity test under each."""
    import os, subprocess, sys
    env = dict(os.environ, UVG266_B200_ROUGH=mode)
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_parity.py"), "-x", "-q",
                        "-k", "test_batched_rough_cost or test_full_size_properties"], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_transform_kernel_variants():
    """Square 16x16 / 32x32 transforms run on the INT8 tensor cores (transform_imma.cu); UVG266_B200_TRANSFORM=fast
    selects the packed-ALU kernel every other shape uses: the same parity tests under it."""
    import os, subprocess, sys
    env = dict(os.environ, UVG266_B200_TRANSFORM="fast")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_parity.py"), "-x", "-q",
                        "-k", "test_dropin_transforms or test_batched_quantize_residual_and_transform or test_transform_full_range"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("bd", BDS)
@pytest.mark.parametrize("n", [16, 32])
def test_transform_full_range_all_type_pairs(bd, n):
    """The tensor-core transform kernels split 16-bit values into byte planes: full-range int16 input (every
    accumulator at its bound), the type pairs of explicit MTS with the zero-out of 32-point DST7 / DCT8, both directions, against
    the oracle's matrix formulation on every TU."""
    O = load_oracle(bd)
    L = U.lib()
    rng = np.random.default_rng(5100 + bd + n)
    nt = 37
    x = rng.integers(-32768, 32768, nt * n * n).astype(np.int16)
    x[: n * n] = 32767; x[n * n: 2 * n * n] = -32768
    x[2 * n * n: 3 * n * n] = np.where(rng.integers(0, 2, n * n) == 1, 32767, -32768)
    # tr_idx -> (type_hor, type_ver) of explicit MTS (dct-generic.c:2501-2557); DCT8 = 1, DST7 = 2
    for tr_idx, (th, tv) in {0: (0, 0), 2: (2, 2), 3: (1, 2), 4: (2, 1), 5: (1, 1)}.items():
        sk = 16 if (n == 32 and th != 0) else 0
        skv = 16 if (n == 32 and tv != 0) else 0
        y = np.zeros_like(x); z = np.zeros_like(x)
        U.check(L.uvg_b200_transform_host(0, bd, n, n, th, tv, sk, skv, ptr(x), ptr(y), nt))
        U.check(L.uvg_b200_transform_host(1, bd, n, n, th, tv, sk, skv, ptr(x), ptr(z), nt))
        for t in range(nt):
            sl = slice(t * n * n, (t + 1) * n * n)
            xx = np.ascontiguousarray(x[sl]); o = np.zeros(n * n, np.int16)
            O.orc_mts_dct(bd, 0, 1, tr_idx, 0, 0, 0, n, n, ptr(xx), ptr(o), 3)
            assert (y[sl] == o).all(), ("fwd", n, th, tv, t)
            O.orc_mts_idct(bd, 0, 1, tr_idx, 0, 0, 0, n, n, ptr(xx), ptr(o), 3)
            assert (z[sl] == o).all(), ("inv", n, th, tv, t)


@pytest.mark.parametrize("bd", BDS)
def test_crc32c_and_bipred_average(bd):
    """The non-distortion leftovers of the picture table: IBC block hash and bi-prediction average,
    drop-ins and batched entry points."""
    import torch
    O = load_oracle(bd)
    L = U.lib()
    O.orc_crc32c_nxn.argtypes = [C.c_int, vp, C.c_uint32]; O.orc_crc32c_nxn.restype = C.c_uint32
    O.orc_bipred_average.argtypes = [vp, C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int]
    rng = np.random.default_rng(2600 + bd)
    hi = (1 << bd) - 1
    PW, PH = 64, 48
    plane = rng.integers(0, hi + 1, (PH, PW)).astype(pix_dtype(bd))
    stream = torch.cuda.current_stream().cuda_stream
    d_plane = torch.from_numpy(plane.view(np.uint8 if bd == 8 else np.int16)).cuda()
    for n in (4, 8):
        f = U.dropin("crc32c_%dx%d" % (n, n), bd)
        for (y, x) in ((0, 0), (3, 5), (PH - n, PW - n)):
            p = plane.ctypes.data + (y * PW + x) * plane.itemsize
            assert f(p, PW) == O.orc_crc32c_nxn(n, p, PW)
        desc = U.Blocks.plane(d_plane.data_ptr(), PW, PW, n, n)
        nb = (PW // n) * (PH // n)
        out = torch.zeros(nb, dtype=torch.int32, device="cuda")
        U.check(L.uvg_b200_crc32c_dev(bd, n, C.byref(desc), out.data_ptr(), nb, stream))
        got = out.cpu().numpy().astype(np.uint32)
        for i in range(nb):
            y, x = (i // (PW // n)) * n, (i % (PW // n)) * n
            assert got[i] == O.orc_crc32c_nxn(n, plane.ctypes.data + (y * PW + x) * plane.itemsize, PW)
    g = U.dropin("bipred_average", bd)
    for (w, h) in ((4, 4), (8, 16), (32, 32), (64, 64)):
        for (i0, i1) in ((0, 0), (1, 1), (0, 1), (1, 0)):
            srcs = [rng.integers(-2000, (hi << (14 - bd)) + 2000, w * h).astype(np.int16) if im else
                    rng.integers(0, hi + 1, w * h).astype(pix_dtype(bd)) for im in (i0, i1)]
            d1 = np.zeros((h, 70), pix_dtype(bd)); d2 = np.zeros((h, 70), pix_dtype(bd))
            O.orc_bipred_average(ptr(d1), 70, ptr(srcs[0]), i0, ptr(srcs[1]), i1, w, h)
            g(ptr(d2), 70, ptr(srcs[0]), i0, ptr(srcs[1]), i1, w, h)
            assert (d1 == d2).all(), (bd, w, h, i0, i1)
    # batched: 10 blocks of 16x16, both sources intermediates
    w = h = 16; nb = 10
    a = rng.integers(-500, (hi << (14 - bd)) + 500, nb * w * h).astype(np.int16); b = rng.integers(-500, (hi << (14 - bd)) + 500, nb * w * h).astype(np.int16)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    dd = torch.zeros(nb * w * h, dtype=torch.uint8 if bd == 8 else torch.int16, device="cuda")
    desc = U.Blocks.dense(dd.data_ptr(), w, h)
    U.check(L.uvg_b200_bipred_average_dev(bd, w, h, C.byref(desc), da.data_ptr(), 1, db.data_ptr(), 1, nb, stream))
    got = dd.cpu().numpy()
    got = got.view(np.uint16) if bd != 8 else got
    for i in range(nb):
        e = np.zeros(w * h, pix_dtype(bd))
        O.orc_bipred_average(ptr(e), w, ptr(a[i * w * h:]), 1, ptr(b[i * w * h:]), 1, w, h)
        assert (got[i * w * h:(i + 1) * w * h] == e).all()
