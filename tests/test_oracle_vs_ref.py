"""Pins the oracle (oracle/uvg_oracle.c) to the UNMODIFIED reference compiled into
oracle/_ref/ (strategy "generic"), on seeded random inputs, 8-bit and 10-bit builds.

CPU only; skipped when oracle/_ref is absent. Covers the functions the reference has no
unit test for: quant, dequant, quantize_residual, ssd, satd_any_size*,
dual, vtm, residual, mts non-square / LFNST zero-out, all 10-bit paths.
"""
import ctypes as C
import numpy as np
import pytest

from _libs import load_oracle, load_ref, ref_available, ptr, pix_dtype, OrcQuantParams

pytestmark = pytest.mark.skipif(not ref_available(8), reason="oracle/_ref not built")

vp, ci, cu = C.c_void_p, C.c_int, C.c_uint
BDS = [8, 10]


def rnd_pix(rng, bd, n, extreme=False):
    hi = (1 << bd)
    if extreme:
        return rng.choice(np.array([0, hi - 1]), n).astype(pix_dtype(bd))
    return rng.integers(0, hi, n).astype(pix_dtype(bd))


@pytest.mark.parametrize("bd", BDS)
def test_picture_pair_functions(bd):
    O, R = load_oracle(bd), load_ref(bd)
    rng = np.random.default_rng(100 + bd)
    for n in (4, 8, 16, 32, 64):
        sad = R.fn("sad_%dx%d" % (n, n), cu, [vp, vp])
        satd = R.fn("satd_%dx%d" % (n, n), cu, [vp, vp])
        sad2 = R.fn("sad_%dx%d_dual" % (n, n), None, [vp, vp, cu, vp])
        satd2 = R.fn("satd_%dx%d_dual" % (n, n), None, [vp, vp, cu, vp])
        for it in range(12):
            a = rnd_pix(rng, bd, n * n, extreme=(it == 0))
            b = rnd_pix(rng, bd, n * n, extreme=(it == 0))
            assert O.orc_sad_nxn(n, ptr(a), ptr(b)) == sad(ptr(a), ptr(b))
            assert O.orc_satd_nxn(n, ptr(a), ptr(b)) == satd(ptr(a), ptr(b))
            if True:
                # pred_buffer slots are 1024 pixels apart for every n; the (unreachable) 64x64 dual reads
                # 4096 pixels from each slot, i.e. overlapping ranges
                preds = rnd_pix(rng, bd, 1024 + max(1024, n * n))
                c1 = np.zeros(2, np.uint32); c2 = np.zeros(2, np.uint32)
                O.orc_sad_nxn_dual(n, ptr(preds), ptr(a), ptr(c1)); sad2(ptr(preds), ptr(a), 2, ptr(c2))
                assert (c1 == c2).all()
                O.orc_satd_nxn_dual(n, ptr(preds), ptr(a), ptr(c1)); satd2(ptr(preds), ptr(a), 2, ptr(c2))
                assert (c1 == c2).all()


@pytest.mark.parametrize("bd", BDS)
def test_picture_strided_functions(bd):
    O, R = load_oracle(bd), load_ref(bd)
    rng = np.random.default_rng(200 + bd)
    reg = R.fn("reg_sad", cu, [vp, vp, ci, ci, cu, cu])
    anys = R.fn("satd_any_size", cu, [ci, ci, vp, ci, vp, ci])
    vtm = R.fn("satd_any_size_vtm", cu, [ci, ci, vp, ci, vp, ci])
    quad = R.fn("satd_any_size_quad", None, [ci, ci, vp, ci, vp, ci, cu, vp, vp])
    ssd = R.fn("pixels_calc_ssd", cu, [vp, vp, ci, ci, ci, ci])
    ver = R.fn("ver_sad", C.c_uint32, [vp, vp, ci, ci, C.c_uint32])
    hor = R.fn("hor_sad", C.c_uint32, [vp, vp, ci, ci] + [C.c_uint32] * 4)
    res = R.fn("generate_residual", None, [vp, vp, vp, ci, ci, ci, ci])
    dims = [(w, h) for w in (4, 8, 12, 16, 24, 32, 48, 64) for h in (4, 8, 12, 16, 32, 64)]
    S1, S2 = 80, 72
    isz = np.dtype(pix_dtype(bd)).itemsize
    for (w, h) in dims:
        a = rnd_pix(rng, bd, S1 * 66, extreme=(w == 64 and h == 64))
        b = rnd_pix(rng, bd, S2 * 66, extreme=(w == 64 and h == 64))
        assert O.orc_reg_sad(ptr(a), ptr(b), w, h, S1, S2) == reg(ptr(a), ptr(b), w, h, S1, S2)
        assert O.orc_satd_any_size(w, h, ptr(a), S1, ptr(b), S2) == anys(w, h, ptr(a), S1, ptr(b), S2)
        assert O.orc_satd_any_size_vtm(w, h, ptr(a), S1, ptr(b), S2) == vtm(w, h, ptr(a), S1, ptr(b), S2)
        assert O.orc_pixels_calc_ssd(ptr(a), ptr(b), S1, S2, w, h) == ssd(ptr(a), ptr(b), S1, S2, w, h)
        assert O.orc_ver_sad(ptr(a), ptr(b), w, h, S1) == ver(ptr(a), ptr(b), w, h, S1)
        for (l, r) in ((0, 0), (3, 0), (0, 2), (w - 1, 0), (0, w - 1)):
            assert O.orc_hor_sad(ptr(a), ptr(b), w, h, S1, S2, l, r) == hor(ptr(a), ptr(b), w, h, S1, S2, l, r)
        r1 = np.zeros(w * h, np.int16); r2 = np.zeros(w * h, np.int16)
        O.orc_generate_residual(ptr(a), ptr(b), ptr(r1), w, h, S1, S2)
        res(ptr(a), ptr(b), ptr(r2), w, h, S1, S2)
        assert (r1 == r2).all()
        # quad: 4 preds with a common stride, literal edge handling included
        P = [rnd_pix(rng, bd, S2 * 66) for _ in range(4)]
        arr = (vp * 4)(*[p.ctypes.data for p in P])
        c1 = np.zeros(4, np.uint32); c2 = np.zeros(4, np.uint32)
        O.orc_satd_any_size_quad(w, h, arr, S2, ptr(a), S1, ptr(c1))
        quad(w, h, arr, S2, ptr(a), S1, 4, ptr(c2), None)
        assert (c1 == c2).all(), (w, h)
    # 2x2-class shapes of xGetHADs
    for (w, h) in ((2, 2), (2, 4), (4, 2), (2, 8), (6, 6), (2, 16)):
        a = rnd_pix(rng, bd, S1 * 20); b = rnd_pix(rng, bd, S2 * 20)
        assert O.orc_satd_any_size_vtm(w, h, ptr(a), S1, ptr(b), S2) == vtm(w, h, ptr(a), S1, ptr(b), S2)


@pytest.mark.parametrize("bd", BDS)
def test_image_calc_sad_random_mvs(bd):
    O, R = load_oracle(bd), load_ref(bd)
    rng = np.random.default_rng(300 + bd)
    W, H = 40, 24
    pic = rnd_pix(rng, bd, W * H); ref = rnd_pix(rng, bd, W * H)
    for _ in range(200):
        bw, bh = int(rng.choice([4, 8, 16])), int(rng.choice([4, 8, 16]))
        px, py = int(rng.integers(0, W - bw + 1)), int(rng.integers(0, H - bh + 1))
        rx, ry = int(rng.integers(-30, W + 30)), int(rng.integers(-30, H + 30))
        got = O.orc_image_calc_sad(ptr(pic), ptr(ref), W, H, W, px, py, rx, ry, bw, bh)
        exp = R.H.refh_image_calc_sad(ptr(pic), ptr(ref), W, H, px, py, rx, ry, bw, bh)
        assert got == exp, (px, py, rx, ry, bw, bh)


def test_tables():
    O, R = load_oracle(8), load_ref(8)
    import os
    L = C.CDLL(os.path.join(os.path.dirname(R.H._name), "libuvg266ref%d.so" % R.bd))
    names = {0: "uvg_g_DCT2P%d", 1: "uvg_g_DCT8P%d", 2: "uvg_g_DST7P%d"}
    for t in (0, 1, 2):
        for n in (2, 4, 8, 16, 32):
            if n == 2 and t != 0:
                continue
            T = np.zeros(n * n, np.int16)
            assert O.orc_tr_matrix(t, n, ptr(T))
            ref = np.array((C.c_int16 * (n * n)).in_dll(L, names[t] % n)[:], np.int16)
            assert (T == ref).all(), (t, n)
    for l2w in range(0, 6):
        for l2h in range(0, 6):
            n = 1 << (l2w + l2h)
            s = np.zeros(n, np.uint32)
            O.orc_scan_table(l2w, l2h, ptr(s))
            p = R.H.refh_scan_table(l2w, l2h)
            assert [p[i] for i in range(n)] == s.tolist(), (l2w, l2h)


@pytest.mark.parametrize("bd", BDS)
def test_square_transforms(bd):
    O, R = load_oracle(bd), load_ref(bd)
    rng = np.random.default_rng(400 + bd)
    sig = [C.c_int8, vp, vp]
    for n in (4, 8, 16, 32):
        f, g = R.fn("dct_%dx%d" % (n, n), None, sig), R.fn("idct_%dx%d" % (n, n), None, sig)
        for amp in (255, 1023, 32767):
            for _ in range(6):
                x = rng.integers(-amp, amp + 1, n * n).astype(np.int16)
                o1 = np.zeros(n * n, np.int16); o2 = np.zeros(n * n, np.int16)
                O.orc_dct_nxn(n, bd, ptr(x), ptr(o1)); f(bd, ptr(x), ptr(o2)); assert (o1 == o2).all()
                O.orc_idct_nxn(n, bd, ptr(x), ptr(o1)); g(bd, ptr(x), ptr(o2)); assert (o1 == o2).all()
    f, g = R.fn("fast_forward_dst_4x4", None, sig), R.fn("fast_inverse_dst_4x4", None, sig)
    for amp in (255, 32767):
        for _ in range(10):
            x = rng.integers(-amp, amp + 1, 16).astype(np.int16)
            o1 = np.zeros(16, np.int16); o2 = np.zeros(16, np.int16)
            O.orc_fast_forward_dst_4x4(bd, ptr(x), ptr(o1)); f(bd, ptr(x), ptr(o2)); assert (o1 == o2).all()
            O.orc_fast_inverse_dst_4x4(bd, ptr(x), ptr(o1)); g(bd, ptr(x), ptr(o2)); assert (o1 == o2).all()


@pytest.mark.parametrize("bd", BDS)
def test_mts(bd):
    O, R = load_oracle(bd), load_ref(bd)
    rng = np.random.default_rng(500 + bd)
    sizes = (1, 2, 4, 8, 16, 32)
    cases = 0
    for w in sizes:
        for h in sizes:
            if w * h < 4 or (w == 2 and h == 2):
                continue  # 2x2 DCT-2 goes to uvg_get_dct_func(2) == NULL in the reference
            for mts_type in (0, 1, 2, 3, 4):
                for cu_type in (1, 2):
                    for tr_idx in (0, 2, 3, 4, 5):
                        for lfnst in (0, 1):
                            for color in (0, 1):
                                isp = 0
                                if (w < 4 or h < 4):
                                    # 1- and 2-point transforms only exist as DCT-2 (dct_table
                                    # NULL entries, dct-generic.c:2484-2488): the encoder only
                                    # reaches them for ISP/chroma blocks with DCT-2.
                                    if color == 0 and mts_type != 0:
                                        continue
                                    if lfnst:
                                        continue
                                hv = (C.c_int * 2)()
                                hv2 = (C.c_int * 2)()
                                O.orc_get_tr_type(w, h, color, cu_type, tr_idx, lfnst, lfnst, isp, mts_type,
                                                  C.byref(hv, 0), C.byref(hv, 4))
                                R.H.refh_get_tr_type(w, h, color, cu_type, tr_idx, lfnst, lfnst, isp, mts_type,
                                                     C.byref(hv2, 0), C.byref(hv2, 4))
                                assert list(hv) == list(hv2)
                                if rng.random() > 0.25:
                                    continue
                                amp = int(rng.choice([255, 1023, 32767]))
                                x = rng.integers(-amp, amp + 1, w * h).astype(np.int16)
                                o1 = np.full(w * h, 77, np.int16); o2 = np.full(w * h, 77, np.int16)
                                for inv in (0, 1):
                                    fn = O.orc_mts_idct if inv else O.orc_mts_dct
                                    fn(bd, color, cu_type, tr_idx, lfnst, lfnst, isp, w, h, ptr(x), ptr(o1), mts_type)
                                    assert R.H.refh_mts(b"generic", inv, bd, color, cu_type, tr_idx, lfnst, lfnst,
                                                        isp, w, h, ptr(x), ptr(o2), mts_type)
                                    assert (o1 == o2).all(), (inv, w, h, mts_type, cu_type, tr_idx, lfnst, color)
                                    cases += 1
    assert cases > 500


QUANT_SHAPES = [(4, 4), (8, 8), (16, 16), (32, 32), (8, 4), (4, 8), (16, 4), (4, 16), (16, 8),
                (8, 16), (32, 8), (8, 32), (32, 16), (16, 32), (2, 8), (8, 2), (2, 16), (32, 4)]


@pytest.mark.parametrize("bd", BDS)
@pytest.mark.parametrize("signhide", [0, 1])
def test_quant_dequant(bd, signhide):
    O = load_oracle(bd)
    R = load_ref(bd, "preset=ultrafast,signhide=%d,rdoq=0" % signhide)
    rng = np.random.default_rng(600 + bd + signhide)
    qmap = R.H.refh_chroma_qp_map()
    n_cases = 0
    for (w, h) in QUANT_SHAPES:
        for qp in (0, 7, 22, 27, 32, 37, 44, 51):
            for slice_i in (0, 1):
                for color in (0, 1, 2):
                    for ts in (0, 1):
                        if rng.random() > 0.35:
                            continue
                        lfnst = int(rng.integers(0, 3)) if (w >= 4 and h >= 4 and rng.random() < 0.3) else 0
                        R.H.refh_set_state(qp, slice_i)
                        p = OrcQuantParams(qp, bd, slice_i, signhide, 0, qmap)
                        amp = int(rng.choice([40, 600, 5000, 32767]))
                        coef = rng.integers(-amp, amp + 1, w * h).astype(np.int16)
                        if rng.random() < 0.3:  # sparse, like real transform output
                            coef[rng.random(w * h) < 0.8] = 0
                        q1 = np.zeros(w * h, np.int16); q2 = np.zeros(w * h, np.int16)
                        O.orc_quant(C.byref(p), ptr(coef), ptr(q1), w, h, color, 1, ts, lfnst)
                        assert R.H.refh_quant(b"generic", ptr(coef), ptr(q2), w, h, color, 0, 1, ts, lfnst)
                        assert (q1 == q2).all(), (w, h, qp, slice_i, color, ts, lfnst)
                        d1 = np.zeros(w * h, np.int16); d2 = np.zeros(w * h, np.int16)
                        O.orc_dequant(C.byref(p), ptr(q1), ptr(d1), w, h, color, 1, ts)
                        assert R.H.refh_dequant(b"generic", ptr(q2), ptr(d2), w, h, color, 1, ts)
                        assert (d1 == d2).all(), (w, h, qp, slice_i, color, ts)
                        n_cases += 1
    assert n_cases > 300


@pytest.mark.parametrize("bd", BDS)
@pytest.mark.parametrize("mts", ["off", "both", "implicit"])
def test_quantize_residual(bd, mts):
    O = load_oracle(bd)
    R = load_ref(bd, "preset=ultrafast,signhide=1,rdoq=0,mts=%s" % mts)
    mts_val = {"off": 0, "both": 3, "implicit": 4}[mts]
    rng = np.random.default_rng(700 + bd + mts_val)
    qmap = R.H.refh_chroma_qp_map()
    shapes = [(4, 4), (8, 8), (16, 16), (32, 32), (16, 8), (8, 32), (4, 16), (32, 4)]
    n_cases = 0
    for (w, h) in shapes:
        for qp in (12, 22, 32, 42):
            for color in (0, 1):
                for tr_idx in ((0, 2, 3, 5) if mts_val == 3 else (0,)):
                    for early in (0, 1):
                        R.H.refh_set_state(qp, 1)
                        p = OrcQuantParams(qp, bd, 1, 1, mts_val, qmap)
                        stride = 64 if color == 0 else 32
                        hi = 1 << bd
                        ref = rng.integers(0, hi, stride * h).astype(pix_dtype(bd))
                        noise = int(rng.choice([2, 12, 60, hi]))
                        pred = np.clip(ref.astype(int) + rng.integers(-noise, noise + 1, stride * h), 0, hi - 1).astype(pix_dtype(bd))
                        rec1 = np.full(stride * h, 7, pix_dtype(bd)); rec2 = rec1.copy()
                        co1 = np.zeros(w * h, np.int16); co2 = np.zeros(w * h, np.int16)
                        hc1 = O.orc_quantize_residual(C.byref(p), 1, tr_idx, w, h, color, 0, stride, stride,
                                                      ptr(ref), ptr(pred), ptr(rec1), ptr(co1), early)
                        hc2 = R.H.refh_quantize_residual(b"generic", 1, tr_idx, 0, w, h, color, 0, stride, stride,
                                                         ptr(ref), ptr(pred), ptr(rec2), ptr(co2), early)
                        assert hc1 == hc2
                        assert (co1 == co2).all(), (w, h, qp, color, tr_idx)
                        assert (rec1 == rec2).all(), (w, h, qp, color, tr_idx, early)
                        n_cases += 1
    assert n_cases >= 128


@pytest.mark.parametrize("bd", BDS)
@pytest.mark.parametrize("signhide", [0, 1])
def test_quantize_residual_transform_skip_and_isp(bd, signhide):
    """The branches the plain test never reaches: use_trskip = 1 (uvg_transformskip, transform.c:223) with and
    without tr_idx == MTS_SKIP - quant only learns about the skip through tr_idx (quant-generic.c:538), so
    (use_trskip 1, tr_idx 0), what chroma transform skip does, keeps the transform shift in q_bits -, tr_idx ==
    MTS_SKIP without use_trskip, and ISP CUs with explicit MTS (DST-7 for 4..16, dct-generic.c:2522-2544)."""
    O = load_oracle(bd)
    R = load_ref(bd, "preset=ultrafast,signhide=%d,rdoq=0,mts=both,transform-skip=1" % signhide)
    rng = np.random.default_rng(750 + bd + signhide)
    qmap = R.H.refh_chroma_qp_map()
    n_cases = 0
    try:
        for (w, h) in [(4, 4), (8, 8), (16, 16), (32, 32), (16, 4), (4, 16), (8, 32), (32, 8)]:
            for qp in (7, 22, 37):
                for color in (0, 1):
                    for (trskip, tr_idx, isp) in ((1, 1, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (0, 0, 2), (0, 2, 1)):
                        for early in (0, 1):
                            R.H.refh_set_state(qp, int(rng.integers(0, 2)))
                            R.H.refh_set_isp_mode(isp)
                            p = OrcQuantParams(qp, bd, R.H.refh_slice_is_intra(), signhide, 3, qmap, isp)
                            stride = 64 if color == 0 else 32
                            hi = 1 << bd
                            ref = rng.integers(0, hi, stride * h).astype(pix_dtype(bd))
                            noise = int(rng.choice([2, 12, 60, hi]))
                            pred = np.clip(ref.astype(int) + rng.integers(-noise, noise + 1, stride * h), 0, hi - 1).astype(pix_dtype(bd))
                            rec1 = np.full(stride * h, 7, pix_dtype(bd)); rec2 = rec1.copy()
                            co1 = np.zeros(w * h, np.int16); co2 = np.zeros(w * h, np.int16)
                            hc1 = O.orc_quantize_residual(C.byref(p), 1, tr_idx, w, h, color, trskip, stride, stride,
                                                          ptr(ref), ptr(pred), ptr(rec1), ptr(co1), early)
                            hc2 = R.H.refh_quantize_residual(b"generic", 1, tr_idx, 0, w, h, color, trskip, stride, stride,
                                                             ptr(ref), ptr(pred), ptr(rec2), ptr(co2), early)
                            assert hc1 == hc2, (w, h, qp, color, trskip, tr_idx, isp)
                            assert (co1 == co2).all(), (w, h, qp, color, trskip, tr_idx, isp)
                            assert (rec1 == rec2).all(), (w, h, qp, color, trskip, tr_idx, isp, early)
                            n_cases += 1
    finally:
        R.H.refh_set_isp_mode(0)
    assert n_cases >= 500


@pytest.mark.parametrize("bd", BDS)
@pytest.mark.parametrize("signhide,mts", [(0, "off"), (1, "off"), (0, "both"), (1, "implicit")])
def test_quant_cbcr_residual(bd, signhide, mts):
    """JCCR (quant-generic.c:241-442): all six joint modes, square and non-square chroma TUs, early skip;
    with MTS enabled the chroma transform stays DCT-2 but goes through mts_dct (transform.c:256-292)."""
    O = load_oracle(bd)
    R = load_ref(bd, "preset=ultrafast,signhide=%d,rdoq=0,jccr=1,mts=%s" % (signhide, mts))
    mts_val = {"off": 0, "both": 3, "implicit": 4}[mts]
    rng = np.random.default_rng(900 + bd + signhide + mts_val)
    qmap = R.H.refh_chroma_qp_map()
    n_cases = 0
    for (w, h) in ((4, 4), (8, 8), (16, 16), (32, 32), (8, 4), (4, 16), (16, 8)):
        for qp in (17, 27, 37):
            for joint in (1, 2, 3):
                for sign in (0, 1):
                    for early in (0, 1):
                        R.H.refh_set_state(qp, 1)
                        p = OrcQuantParams(qp, bd, 1, signhide, mts_val, qmap)
                        stride = 32
                        hi = 1 << bd
                        noise = int(rng.choice([2, 12, 60, hi]))
                        planes = []
                        for _ in range(2):
                            ref = rng.integers(0, hi, stride * h).astype(pix_dtype(bd))
                            pred = np.clip(ref.astype(int) + rng.integers(-noise, noise + 1, stride * h), 0, hi - 1).astype(pix_dtype(bd))
                            planes.append((ref, pred))
                        (ur, up), (vr, vp_) = planes
                        rec = [np.full(stride * h, 7, pix_dtype(bd)) for _ in range(4)]
                        co1 = np.zeros(w * h, np.int16); co2 = np.zeros(w * h, np.int16)
                        r1 = O.orc_quant_cbcr_residual(C.byref(p), 1, 0, w, h, joint, sign, stride, stride, ptr(ur), ptr(vr),
                                                       ptr(up), ptr(vp_), ptr(rec[0]), ptr(rec[1]), ptr(co1), early)
                        r2 = R.H.refh_quant_cbcr_residual(b"generic", 1, 0, joint, sign, w, h, stride, stride, ptr(ur), ptr(vr),
                                                          ptr(up), ptr(vp_), ptr(rec[2]), ptr(rec[3]), ptr(co2), early)
                        assert r1 == r2, (w, h, qp, joint, sign, early)
                        assert (co1 == co2).all(), (w, h, qp, joint, sign)
                        assert (rec[0] == rec[2]).all() and (rec[1] == rec[3]).all(), (w, h, qp, joint, sign, early)
                        n_cases += 1
    assert n_cases >= 250


def test_coeff_costs():
    O, R = load_oracle(8), load_ref(8)
    rng = np.random.default_rng(800)
    fcc = R.fn("fast_coeff_cost", C.c_uint32, [vp, ci, ci, C.c_uint64])
    cas = R.fn("coeff_abs_sum", C.c_uint32, [vp, C.c_size_t])
    for n in (4, 8, 16, 32):
        for _ in range(10):
            c = rng.integers(-6, 7, n * n).astype(np.int16)
            wts = int(rng.integers(0, 1 << 62))
            assert O.orc_fast_coeff_cost(ptr(c), n, n, wts) == fcc(ptr(c), n, n, wts)
            big = rng.integers(-32768, 32768, n * n).astype(np.int16)
            assert O.orc_coeff_abs_sum(ptr(big), big.size) == cas(ptr(big), big.size)
