"""LFNST (first piece): the oracle's restatement of uvg_fwd_lfnst / uvg_inv_lfnst
(transform.c:878-1200) and of quantize_residual with cfg.lfnst, pinned on the compiled reference.
The trained kernels are not in this repository: the oracle receives their addresses from the
reference library (exported data symbols uvg_lfnst_8x8 / uvg_lfnst_4x4)."""
import ctypes as C

import numpy as np
import pytest

from _libs import OrcQuantParams, load_oracle, load_ref, pix_dtype, ptr, ref_available

pytestmark = pytest.mark.skipif(not ref_available(8), reason="oracle/_ref not built")
BDS = [8, 10]


class OrcLfnstInfo(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("lfnst_idx", "cr_lfnst_idx", "intra_mode", "mode_chroma", "cu_log2_w",
                                       "cu_log2_h", "tree_type", "luma_mode")]


def _bind(O, R):
    i, vp = C.c_int, C.c_void_p
    R.H.refh_lfnst_table.argtypes = [i]; R.H.refh_lfnst_table.restype = vp
    R.H.refh_lfnst.argtypes = [i] * 12 + [vp]; R.H.refh_lfnst.restype = None
    R.H.refh_quantize_residual_lfnst.argtypes = [C.c_char_p] + [i] * 15 + [vp] * 4 + [i]
    R.H.refh_quantize_residual_lfnst.restype = i
    O.orc_set_lfnst_tables.argtypes = [vp, vp]
    O.orc_fwd_lfnst.argtypes = [i] * 6 + [vp]; O.orc_inv_lfnst.argtypes = [i] * 6 + [vp]
    O.orc_quantize_residual_lfnst.argtypes = [vp, vp, i, i, i, i, i, i, i, vp, vp, vp, vp, i]
    O.orc_quantize_residual_lfnst.restype = i
    O.orc_set_lfnst_tables(R.H.refh_lfnst_table(1), R.H.refh_lfnst_table(0))


SHAPES = [(4, 4), (8, 8), (16, 16), (32, 32), (4, 8), (8, 4), (4, 16), (16, 4), (8, 16), (16, 8), (8, 32), (32, 8)]


@pytest.mark.parametrize("bd", BDS)
def test_fwd_and_inv_lfnst(bd):
    """Every intra mode (incl. the ones the wide-angle correction moves outside 2..66), both kernels of every
    set, all TU shapes LFNST applies to, luma and separate-tree chroma."""
    O, R = load_oracle(bd), load_ref(bd, "preset=ultrafast,lfnst=1")
    _bind(O, R)
    rng = np.random.default_rng(2100 + bd)
    n = 0
    for (w, h) in SHAPES:
        l2w, l2h = w.bit_length() - 1, h.bit_length() - 1
        for mode in list(range(0, 67)):
            for lfnst_idx in (1, 2):
                for color, tree in ((0, 0), (1, 2)):
                    c = rng.integers(-2000, 2001, w * h).astype(np.int16)
                    if mode % 7 == 0:
                        c = rng.integers(-32768, 32768, w * h).astype(np.int16)       # wrap / cast behaviour
                    for inverse in (0, 1):
                        a, b = c.copy(), c.copy()
                        (O.orc_inv_lfnst if inverse else O.orc_fwd_lfnst)(w, h, lfnst_idx, mode, l2w, l2h, ptr(a))
                        R.H.refh_lfnst(inverse, w, h, color, lfnst_idx, 0, mode, mode, l2w, l2h, tree, 0, ptr(b))
                        assert (a == b).all(), (w, h, mode, lfnst_idx, color, inverse)
                        n += 1
    assert n > 6000


@pytest.mark.parametrize("bd", BDS)
@pytest.mark.parametrize("mts", ["off", "intra"])
def test_quantize_residual_with_lfnst(bd, mts):
    """The composite with cfg.lfnst (quant-generic.c:497-566): LFNST between transform and quant and between
    dequant and inverse transform, the quantiser's LFNST zero-out, and (with MTS on) mts_dct's zero-out."""
    O = load_oracle(bd)
    R = load_ref(bd, "preset=ultrafast,signhide=0,rdoq=0,lfnst=1,mts=%s" % mts)
    _bind(O, R)
    mts_val = {"off": 0, "intra": 1}[mts]
    rng = np.random.default_rng(2200 + bd + mts_val)
    qmap = R.H.refh_chroma_qp_map()
    hi = 1 << bd
    n = 0
    for (w, h) in SHAPES:
        l2w, l2h = w.bit_length() - 1, h.bit_length() - 1
        for qp in (17, 27, 37):
            for mode in (0, 1, 2, 18, 34, 35, 50, 66):
                for lfnst_idx in (0, 1, 2):
                    for color, tree in ((0, 0), (1, 2), (1, 0)):
                        R.H.refh_set_state(qp, 1)
                        p = OrcQuantParams(qp, bd, 1, 0, mts_val, qmap)
                        li = OrcLfnstInfo(lfnst_idx, lfnst_idx, mode, mode, l2w, l2h, tree, 0)
                        stride = 64
                        ref = rng.integers(0, hi, stride * h).astype(pix_dtype(bd))
                        noise = int(rng.choice([4, 20, 80]))
                        pred = np.clip(ref.astype(int) + rng.integers(-noise, noise + 1, stride * h), 0, hi - 1).astype(pix_dtype(bd))
                        rec1 = np.full(stride * h, 7, pix_dtype(bd)); rec2 = rec1.copy()
                        co1 = np.zeros(w * h, np.int16); co2 = np.zeros(w * h, np.int16)
                        h1 = O.orc_quantize_residual_lfnst(C.byref(p), C.byref(li), 1, 0, w, h, color, stride, stride, ptr(ref),
                                                           ptr(pred), ptr(rec1), ptr(co1), 0)
                        h2 = R.H.refh_quantize_residual_lfnst(b"generic", 1, 0, lfnst_idx, lfnst_idx, mode, mode, l2w, l2h, tree, 0,
                                                              w, h, color, stride, stride, ptr(ref), ptr(pred), ptr(rec2), ptr(co2), 0)
                        assert h1 == h2, (w, h, qp, mode, lfnst_idx, color, tree)
                        assert (co1 == co2).all(), (w, h, qp, mode, lfnst_idx, color, tree)
                        assert (rec1 == rec2).all(), (w, h, qp, mode, lfnst_idx, color, tree)
                        n += 1
    assert n > 2000
