"""GPU parity, next row 8(f)-2: batched intra prediction and the fused predict+score rough-search
kernel against the oracle (orc_intra_predict, pinned to the reference's uvg_intra_predict by
tests/test_oracle_intra_vs_ref.py; costs = min(SATD, 2*SAD) like search_intra.c:149-163)."""
import ctypes as C

import numpy as np
import pytest

import uvg266_b200 as U
from _libs import load_oracle, ptr, pix_dtype

pytestmark = pytest.mark.gpu
ci, vp = C.c_int, C.c_void_p


@pytest.fixture(scope="module", autouse=True)
def _init():
    U.init(0)


def make_refs(rng, bd, n, nb, kind):
    """[nb][2][2n+4] reference rows (top, left); the 3 trailing pad samples are garbage on purpose."""
    hi = (1 << bd) - 1
    gp = 2 * n + 4
    if kind == "random":
        r = rng.integers(0, hi + 1, (nb, 2, gp))
    elif kind == "smooth":
        base = rng.integers(0, hi + 1, (nb, 1, 1))
        r = np.clip(base + np.cumsum(rng.integers(-3, 4, (nb, 2, gp)), axis=2), 0, hi)
    else:
        r = rng.choice([0, hi], (nb, 2, gp))
    r[:, 1, 0] = r[:, 0, 0]
    return np.ascontiguousarray(r.astype(pix_dtype(bd)))


def oracle_pred(O, bd, n, refs, b, mode):
    top = np.ascontiguousarray(refs[b, 0, :2 * n + 1]); left = np.ascontiguousarray(refs[b, 1, :2 * n + 1])
    d = np.zeros(n * n, pix_dtype(bd))
    O.orc_intra_predict(n, n, mode, ptr(top), ptr(left), ptr(d))
    return d


@pytest.mark.parametrize("bd", [8, 10])
@pytest.mark.parametrize("n", [4, 8, 16, 32])
def test_intra_predict_all_modes(bd, n):
    O = load_oracle(bd)
    O.orc_intra_predict.argtypes = [ci, ci, ci, vp, vp, vp]
    L = U.lib()
    rng = np.random.default_rng(5100 + bd + n)
    modes = np.arange(67, dtype=np.int8)
    for kind in ("random", "smooth", "extreme"):
        nb = 6
        refs = make_refs(rng, bd, n, nb, kind)
        preds = np.zeros((nb, 67, n * n), pix_dtype(bd))
        U.check(L.uvg_b200_intra_predict_host(bd, n, ptr(refs), ptr(modes), 67, ptr(preds), nb))
        for b in range(nb):
            for m in range(67):
                e = oracle_pred(O, bd, n, refs, b, m)
                assert (preds[b, m] == e).all(), (bd, n, kind, b, m)


@pytest.mark.parametrize("n,k", [(8, 23), (16, 24), (32, 67), (8, 67)])
def test_intra_rough_fused(n, k):
    bd = 8
    O = load_oracle(bd)
    O.orc_intra_predict.argtypes = [ci, ci, ci, vp, vp, vp]
    L = U.lib()
    rng = np.random.default_rng(5200 + n + k)
    W, H = 128, 64
    nb = (W // n) * (H // n)
    plane = np.clip(128 + np.cumsum(rng.integers(-6, 7, (H, W)), axis=1), 0, 255).astype(np.uint8)
    refs = make_refs(rng, bd, n, nb, "smooth")
    modes = (np.arange(67, dtype=np.int8) if k == 67 else
             np.array(sorted(rng.choice(67, k, replace=False)), dtype=np.int8))
    out = np.zeros(nb * k, np.uint32)
    desc = U.Blocks.plane(plane.ctypes.data, W, W, n, n)
    U.check(L.uvg_b200_intra_rough_host(bd, n, C.byref(desc), plane.nbytes, ptr(refs), ptr(modes), k, ptr(out), nb))
    for b in range(0, nb, max(1, nb // 12)):
        y0, x0 = (b // (W // n)) * n, (b % (W // n)) * n
        orig = np.ascontiguousarray(plane[y0:y0 + n, x0:x0 + n]).ravel()
        for j, m in enumerate(modes):
            pred = oracle_pred(O, bd, n, refs, b, int(m))
            pb = np.zeros(2048, np.uint8); pb[:n * n] = pred; pb[1024:1024 + n * n] = pred
            s = np.zeros(2, np.uint32); t = np.zeros(2, np.uint32)
            O.orc_satd_nxn_dual(n, ptr(pb), ptr(orig), ptr(s)); O.orc_sad_nxn_dual(n, ptr(pb), ptr(orig), ptr(t))
            assert out[b * k + j] == min(int(s[0]), 2 * int(t[0])), (n, b, int(m))


def test_intra_errors():
    L = U.lib()
    modes = np.array([0, 1, 67], dtype=np.int8)
    buf = np.zeros(4096, np.uint8)
    assert L.uvg_b200_intra_predict_host(8, 8, ptr(buf), ptr(modes), 3, ptr(buf), 1) != 0
    assert b"mode" in L.uvg_b200_last_error()
    assert L.uvg_b200_intra_predict_host(8, 12, ptr(buf), ptr(modes), 2, ptr(buf), 1) != 0
    desc = U.Blocks.dense(buf.ctypes.data, 8, 8)
    assert L.uvg_b200_intra_rough_host(10, 8, C.byref(desc), 64, ptr(buf), ptr(modes), 2, ptr(buf), 1) != 0


@pytest.mark.parametrize("bd", [8, 10])
def test_intra_leaf_dropins(bd):
    """The strategy-leaf drop-ins (angular_pred / intra_pred_planar / pdpc_planar_dc signatures) with every
    argument the encoder passes: non-square, wide angles, chroma, multi_ref_idx, ISP."""
    O = load_oracle(bd)
    O.orc_angular_pred_leaf.argtypes = [ci, ci, ci, ci, vp, vp, vp, ci, ci]
    O.orc_intra_pred_planar.argtypes = [ci, ci, vp, vp, vp]
    O.orc_pdpc_planar_dc.argtypes = [ci, ci, vp, vp, vp]
    L = U.lib()
    L.uvg_b200_set_dropin_bitdepth(10)
    suf = U.suffix(bd)
    ang, pla, pdp = (getattr(L, "uvg_b200_%s_%s" % (n, suf)) for n in ("angular_pred", "intra_pred_planar", "pdpc_planar_dc"))
    rng = np.random.default_rng(5300 + bd)
    hi = (1 << bd) - 1
    RL = 2 * 128 + 3 + 33 * 3
    # ISP predicts strips of a CU with the CU's (uncorrected) mode: every mode on every shape, sizes down to 2
    for (w, h) in [(4, 4), (8, 8), (16, 16), (32, 32), (16, 4), (4, 16), (32, 8), (8, 32), (64, 64), (8, 2), (2, 8),
                   (16, 2), (2, 16), (32, 2), (2, 32), (4, 2), (8, 4), (4, 8), (32, 4), (4, 32), (16, 8), (32, 16)]:
        above = rng.integers(0, hi + 1, RL).astype(pix_dtype(bd)); left = rng.integers(0, hi + 1, RL).astype(pix_dtype(bd))
        left[0] = above[0]
        ratio = abs(int(np.log2(w)) - int(np.log2(h)))
        extra = 0 if w == h else (6 if ratio == 1 else 10 if ratio == 2 else 12 if ratio == 3 else 14)
        modes = list(range(2, 67))
        if w > h: modes += list(range(67, 67 + extra))
        if h > w: modes += list(range(2 - extra, 2))
        for mode in modes[::2] if w * h >= 256 else modes:
            for (ch, mrl, isp) in ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 2, 0), (0, 0, 1), (0, 0, 2)):
                if max(w, h) > 32 and ch == 0:
                    continue
                d1 = np.zeros(w * h, pix_dtype(bd)); d2 = np.zeros(w * h, pix_dtype(bd))
                O.orc_angular_pred_leaf(w, h, mode, ch, ptr(above), ptr(left), ptr(d1), mrl, isp)
                ang(w, h, mode, ch, ptr(above), ptr(left), ptr(d2), mrl, isp)
                assert (d1 == d2).all(), (bd, w, h, mode, ch, mrl, isp)
        if max(w, h) <= 32:
            d1 = np.zeros(w * h, pix_dtype(bd)); d2 = np.zeros(w * h, pix_dtype(bd))
            O.orc_intra_pred_planar(w, h, ptr(above), ptr(left), ptr(d1)); pla(w, h, ptr(above), ptr(left), ptr(d2))
            assert (d1 == d2).all(), ("planar", bd, w, h)
        if max(w, h) <= 32 and min(w, h) >= 4:
            d1[:] = rng.integers(0, hi + 1, w * h); d2[:] = d1
            O.orc_pdpc_planar_dc(w, h, ptr(above), ptr(left), ptr(d1)); pdp(1, w, h, ptr(above), ptr(left), ptr(d2))
            assert (d1 == d2).all(), ("pdpc", bd, w, h)
