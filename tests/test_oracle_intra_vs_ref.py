"""Pins the intra-prediction restatement (oracle: orc_intra_predict and its leaves; next row
8(f)-2) to the unmodified reference's uvg_intra_predict (intra.c:1372 -> intra_predict_regular
-> angular / planar / DC + PDPC, generic strategies), all 67 modes, square and non-square luma
blocks, 8-bit and 10-bit. CPU only."""
import ctypes as C

import numpy as np
import pytest

from _libs import load_oracle, load_ref, ref_available, ptr, pix_dtype

pytestmark = pytest.mark.skipif(not ref_available(8), reason="oracle/_ref not built")
vp, ci = C.c_void_p, C.c_int


def refs_for(rng, bd, w, h, kind):
    hi = (1 << bd) - 1
    nt, nl = 2 * w + 1, 2 * h + 1
    if kind == "random":
        top = rng.integers(0, hi + 1, nt); left = rng.integers(0, hi + 1, nl)
    elif kind == "smooth":
        base = rng.integers(0, hi + 1)
        top = np.clip(base + np.cumsum(rng.integers(-3, 4, nt)), 0, hi)
        left = np.clip(base + np.cumsum(rng.integers(-3, 4, nl)), 0, hi)
    else:  # extreme
        top = rng.choice([0, hi], nt); left = rng.choice([0, hi], nl)
    left[0] = top[0]     # shared top-left corner sample
    return top.astype(pix_dtype(bd)), left.astype(pix_dtype(bd))


@pytest.mark.parametrize("bd", [8, 10])
def test_intra_predict_all_modes(bd):
    O, R = load_oracle(bd), load_ref(bd)
    O.orc_intra_predict.argtypes = [ci, ci, ci, vp, vp, vp]
    R.H.refh_intra_predict.argtypes = [ci, ci, ci, vp, vp, vp]
    rng = np.random.default_rng(4200 + bd)
    shapes = [(4, 4), (8, 8), (16, 16), (32, 32), (8, 4), (4, 8), (16, 4), (4, 16), (16, 8), (8, 16), (32, 8),
              (8, 32), (32, 16), (16, 32)]
    n = 0
    for (w, h) in shapes:
        for kind in ("random", "smooth", "extreme"):
            top, left = refs_for(rng, bd, w, h, kind)
            for mode in range(67):
                d1 = np.zeros(w * h, pix_dtype(bd)); d2 = np.zeros(w * h, pix_dtype(bd))
                O.orc_intra_predict(w, h, mode, ptr(top), ptr(left), ptr(d1))
                R.H.refh_intra_predict(w, h, mode, ptr(top), ptr(left), ptr(d2))
                assert (d1 == d2).all(), (bd, w, h, kind, mode)
                n += 1
    assert n == len(shapes) * 3 * 67


def test_intra_rough_batch_matches_oracle():
    """The CPU-baseline leg of tools/intra_bench.py (refh_intra_rough_batch) against the oracle."""
    bd = 8
    O, R = load_oracle(bd), load_ref(bd)
    O.orc_intra_predict.argtypes = [ci, ci, ci, vp, vp, vp]
    R.H.refh_intra_rough_batch.argtypes = [ci, vp, ci, vp, vp, C.c_long, vp]
    rng = np.random.default_rng(4300)
    for n in (8, 16, 32):
        nb, gp = 5, 2 * n + 4
        refs = rng.integers(0, 256, (nb, 2, gp)).astype(np.uint8)
        refs[:, 1, 0] = refs[:, 0, 0]
        orig = rng.integers(0, 256, (nb, n * n)).astype(np.uint8)
        modes = np.arange(67, dtype=np.int8)
        costs = np.zeros(nb * 67, np.uint32)
        R.H.refh_intra_rough_batch(n, ptr(modes), 67, ptr(refs), ptr(orig), nb, ptr(costs))
        for b in range(nb):
            top = np.ascontiguousarray(refs[b, 0, :2 * n + 1]); left = np.ascontiguousarray(refs[b, 1, :2 * n + 1])
            for m in range(67):
                d = np.zeros(n * n, np.uint8)
                O.orc_intra_predict(n, n, m, ptr(top), ptr(left), ptr(d))
                e = min(O.orc_satd_nxn(n, ptr(orig[b]), ptr(d)), 2 * O.orc_sad_nxn(n, ptr(orig[b]), ptr(d)))
                assert costs[b * 67 + m] == e, (n, b, m)


@pytest.mark.parametrize("bd", [8, 10])
def test_intra_leaves_all_arguments(bd):
    """The strategy leaves with every argument the encoder can pass: wide-angle corrected modes
    (-14..80), chroma (2-tap filter), multi_ref_idx 1..2 (luma), ISP, non-square blocks."""
    O, R = load_oracle(bd), load_ref(bd)
    L = R.H.refh_intra_ref_len()
    O.orc_angular_pred_leaf.argtypes = [ci, ci, ci, ci, vp, vp, vp, ci, ci]
    R.H.refh_angular_pred_leaf.argtypes = [ci, ci, ci, ci, vp, vp, vp, ci, ci]
    O.orc_intra_pred_planar.argtypes = [ci, ci, vp, vp, vp]
    R.H.refh_intra_pred_planar_leaf.argtypes = [ci, ci, ci, vp, vp, vp]
    O.orc_pdpc_planar_dc.argtypes = [ci, ci, vp, vp, vp]
    R.H.refh_pdpc_planar_dc_leaf.argtypes = [ci, ci, ci, ci, vp, vp, vp]
    rng = np.random.default_rng(4400 + bd)
    hi = (1 << bd) - 1
    shapes = [(4, 4), (8, 8), (16, 16), (32, 32), (8, 4), (4, 8), (16, 4), (4, 16), (32, 8), (8, 32), (16, 32), (64, 64)]
    n = 0
    for (w, h) in shapes:
        above = rng.integers(0, hi + 1, L).astype(pix_dtype(bd)); left = rng.integers(0, hi + 1, L).astype(pix_dtype(bd))
        left[0] = above[0]
        # legal wide-angle range for this shape (uvg_wide_angle_correction, intra.c:637-658)
        ratio = abs(int(np.log2(w)) - int(np.log2(h)))
        extra = 0 if w == h else (6 if ratio == 1 else 10 if ratio == 2 else 12 if ratio == 3 else 14)
        modes = list(range(2, 67))
        if w > h: modes = [m for m in modes if m >= 2 + extra] + list(range(67, 67 + extra))
        if h > w: modes = [m for m in modes if m <= 66 - extra] + list(range(2 - extra, 2))
        for mode in modes:
            for (ch, mrl, isp) in ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 2, 0), (0, 0, 1)):
                if max(w, h) > 32 and ch == 0:
                    continue                    # 64-wide luma blocks are never predicted in one piece
                d1 = np.zeros(w * h, pix_dtype(bd)); d2 = np.zeros(w * h, pix_dtype(bd))
                O.orc_angular_pred_leaf(w, h, mode, ch, ptr(above), ptr(left), ptr(d1), mrl, isp)
                R.H.refh_angular_pred_leaf(w, h, mode, ch, ptr(above), ptr(left), ptr(d2), mrl, isp)
                assert (d1 == d2).all(), (bd, w, h, mode, ch, mrl, isp)
                n += 1
        if max(w, h) <= 32:
            d1 = np.zeros(w * h, pix_dtype(bd)); d2 = np.zeros(w * h, pix_dtype(bd))
            O.orc_intra_pred_planar(w, h, ptr(above), ptr(left), ptr(d1))
            R.H.refh_intra_pred_planar_leaf(w, h, 0, ptr(above), ptr(left), ptr(d2))
            assert (d1 == d2).all(), ("planar", bd, w, h)
            if w >= 4 and h >= 4:
                d1[:] = rng.integers(0, hi + 1, w * h); d2[:] = d1
                O.orc_pdpc_planar_dc(w, h, ptr(above), ptr(left), ptr(d1))
                R.H.refh_pdpc_planar_dc_leaf(1, w, h, 0, ptr(above), ptr(left), ptr(d2))
                assert (d1 == d2).all(), ("pdpc", bd, w, h)
    assert n > 2000
