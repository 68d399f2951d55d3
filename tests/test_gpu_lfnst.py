"""GPU parity of LFNST inside the fused quantize_residual kernel (row 8(f)-3, stage 1): residual -> transform (with
the LFNST zero-out) -> forward LFNST -> quant (first 8 / 16 scan positions) -> dequant -> inverse LFNST -> inverse
transform -> reconstruction in ONE launch, against orc_quantize_residual_lfnst, which tests/test_oracle_lfnst_vs_ref.py
pins on the unmodified reference. The trained kernels are taken from the compiled reference at run time
(exported data uvg_lfnst_8x8 / uvg_lfnst_4x4), exactly as the strategy glue does inside the encoder process."""
import ctypes as C

import numpy as np
import pytest

import uvg266_b200 as U
from _libs import OrcQuantParams, load_oracle, load_ref, pix_dtype, ptr, ref_available
from test_oracle_lfnst_vs_ref import OrcLfnstInfo, SHAPES, _bind

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_available(8), reason="oracle/_ref (source of the LFNST kernels) not built")]
BDS = [8, 10]


def lfnst_cfg(qp, bd, mts_val, qlist, li, color, w, h, tr_idx=0):
    """The flattening the glue does (strategies_cuda.c quantize_residual_cuda): quant-generic.c:503-509 + transform.c:979-1014"""
    cfg = U.QuantCfg.make(qp, bd, 1, 0, mts_val, qlist)
    lfnst_index = li.lfnst_idx if (li.tree_type != 2 or color == 0) else li.cr_lfnst_idx
    sep_tree = li.cu_log2_w + li.cu_log2_h < 6 or li.tree_type != 0
    mode = li.intra_mode if color == 0 else li.mode_chroma
    if 81 <= mode <= 83:
        mode = li.luma_mode
    cfg.cu_lfnst_idx, cfg.cu_cr_lfnst_idx, cfg.lfnst_index = li.lfnst_idx, li.cr_lfnst_idx, lfnst_index
    cfg.lfnst_apply = int(bool(lfnst_index) and not (tr_idx == 1 and color == 0) and (color == 0 or sep_tree))
    cfg.lfnst_mode = mode
    cfg.lfnst_wa_log2_w = li.cu_log2_w if color == 0 else w.bit_length() - 1
    cfg.lfnst_wa_log2_h = li.cu_log2_h if color == 0 else h.bit_length() - 1
    return cfg


@pytest.fixture(scope="module", autouse=True)
def _init():
    U.init(0)
    U.lib().uvg_b200_set_dropin_bitdepth(10)


@pytest.mark.parametrize("bd", BDS)
@pytest.mark.parametrize("mts", ["off", "intra"])
def test_dropin_quantize_residual_with_lfnst(bd, mts):
    O = load_oracle(bd)
    R = load_ref(bd, "preset=ultrafast,signhide=0,rdoq=0,lfnst=1,mts=%s" % mts)
    _bind(O, R)
    L = U.lib()
    R.H.refh_lfnst_table.restype = C.c_void_p
    U.check(L.uvg_b200_set_lfnst_tables(R.H.refh_lfnst_table(1), R.H.refh_lfnst_table(0)))
    mts_val = {"off": 0, "intra": 1}[mts]
    rng = np.random.default_rng(2300 + bd + mts_val)
    qp_ptr = R.H.refh_chroma_qp_map()
    qlist = [(C.c_int8 * 64).from_address(qp_ptr)[i] for i in range(64)]
    f = getattr(L, "uvg_b200_quantize_residual_%s" % U.suffix(bd))
    hi = 1 << bd
    n = 0
    for (w, h) in SHAPES:
        l2w, l2h = w.bit_length() - 1, h.bit_length() - 1
        for qp in (17, 32):
            for mode in (0, 1, 2, 18, 34, 35, 50, 66):
                for lfnst_idx in (0, 1, 2):
                    for color, tree in ((0, 0), (1, 2), (1, 0)):
                        # non-square CUs exercise the wide-angle remap through the CU dims
                        p = OrcQuantParams(qp, bd, 1, 0, mts_val, qp_ptr)
                        li = OrcLfnstInfo(lfnst_idx, lfnst_idx, mode, mode, l2w, l2h, tree, 0)
                        cfg = lfnst_cfg(qp, bd, mts_val, qlist, li, color, w, h)
                        stride = 64
                        ref = rng.integers(0, hi, stride * h).astype(pix_dtype(bd))
                        noise = int(rng.choice([4, 20, 80]))
                        pred = np.clip(ref.astype(int) + rng.integers(-noise, noise + 1, stride * h), 0, hi - 1).astype(pix_dtype(bd))
                        early = n & 1 if n % 5 == 0 else 0
                        rec1 = np.full(stride * h, 7, pix_dtype(bd)); rec2 = rec1.copy()
                        co1 = np.zeros(w * h, np.int16); co2 = np.zeros(w * h, np.int16)
                        h1 = f(C.byref(cfg), 1, 0, w, h, color, 0, stride, stride, ptr(ref), ptr(pred), ptr(rec1), ptr(co1), early)
                        h2 = O.orc_quantize_residual_lfnst(C.byref(p), C.byref(li), 1, 0, w, h, color, stride, stride, ptr(ref),
                                                           ptr(pred), ptr(rec2), ptr(co2), early)
                        assert h1 == h2, (w, h, qp, mode, lfnst_idx, color, tree)
                        assert (co1 == co2).all(), (w, h, qp, mode, lfnst_idx, color, tree)
                        assert (rec1 == rec2).all(), (w, h, qp, mode, lfnst_idx, color, tree)
                        n += 1
    assert n > 1500


@pytest.mark.parametrize("bd", BDS)
def test_batched_quantize_residual_with_lfnst(bd):
    """Many TUs per launch (the shape config 4 of BASELINE.json produces: 10-bit / 8-bit, MTS + LFNST)."""
    O = load_oracle(bd)
    R = load_ref(bd, "preset=ultrafast,signhide=0,rdoq=0,lfnst=1,mts=intra")
    _bind(O, R)
    L = U.lib()
    R.H.refh_lfnst_table.restype = C.c_void_p
    U.check(L.uvg_b200_set_lfnst_tables(R.H.refh_lfnst_table(1), R.H.refh_lfnst_table(0)))
    rng = np.random.default_rng(2400 + bd)
    hi = 1 << bd
    dt = pix_dtype(bd)
    for (n, nt) in ((4, 130), (8, 67), (16, 21), (32, 5)):
        for (mode, lfnst_idx, tr_idx) in ((0, 1, 0), (40, 2, 0), (18, 1, 0), (66, 2, 0)):
            l2 = n.bit_length() - 1
            li = OrcLfnstInfo(lfnst_idx, 0, mode, mode, l2, l2, 0, 0)
            cfg = lfnst_cfg(27, bd, 1, None, li, 0, n, n, tr_idx)
            ref = rng.integers(0, hi, nt * n * n).astype(dt)
            pred = np.clip(ref.astype(int) + rng.integers(-40, 41, nt * n * n), 0, hi - 1).astype(dt)
            pred[: n * n * 2] = ref[: n * n * 2]
            rec = np.zeros(nt * n * n, dt)
            coeff = np.zeros(nt * n * n, np.int16); has = np.zeros(nt, np.uint8)
            ssd = np.zeros(nt, np.uint32)
            Rr, P, Rc = (U.Blocks.dense(a.ctypes.data, n, n) for a in (ref, pred, rec))
            U.check(L.uvg_b200_quantize_residual_host(C.byref(cfg), 1, tr_idx, n, n, 0, 0, 0, C.byref(Rr), ref.nbytes, C.byref(P),
                                                      pred.nbytes, C.byref(Rc), rec.nbytes, ptr(coeff), ptr(has), ptr(ssd), None, 0, nt))
            p = OrcQuantParams(27, bd, 1, 0, 1, None)
            for t in range(nt):
                sl = slice(t * n * n, (t + 1) * n * n)
                r2 = np.zeros(n * n, dt); c2 = np.zeros(n * n, np.int16)
                rr = np.ascontiguousarray(ref[sl]); pp = np.ascontiguousarray(pred[sl])
                h2 = O.orc_quantize_residual_lfnst(C.byref(p), C.byref(li), 1, tr_idx, n, n, 0, n, n, ptr(rr), ptr(pp), ptr(r2), ptr(c2), 0)
                assert has[t] == h2, (n, mode, t)
                assert (coeff[sl] == c2).all() and (rec[sl] == r2).all(), (n, mode, lfnst_idx, t)
                assert ssd[t] == O.orc_pixels_calc_ssd(ptr(rr), ptr(r2), n, n, n, n)
