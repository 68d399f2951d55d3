"""Pins the MC sample-interpolation restatement (next row 8(f)-1: orc_sample_quarterpel_luma(_hi),
orc_sample_octpel_chroma(_hi)) to the unmodified reference's generic strategies
(generic/ipol-generic.c:134-210, :681-759): every fractional phase pair, several block sizes,
8-bit and 10-bit, random and extreme windows. CPU only."""
import ctypes as C

import numpy as np
import pytest

from _libs import load_oracle, load_ref, ref_available, ptr, pix_dtype

pytestmark = pytest.mark.skipif(not ref_available(8), reason="oracle/_ref not built")
vp, ci = C.c_void_p, C.c_int
KINDS = [(0, "orc_sample_quarterpel_luma", 8, 16, False), (1, "orc_sample_quarterpel_luma_hi", 8, 16, True),
         (2, "orc_sample_octpel_chroma", 4, 32, False), (3, "orc_sample_octpel_chroma_hi", 4, 32, True)]


@pytest.mark.parametrize("bd", [8, 10])
def test_sample_interpolation_all_phases(bd):
    O, R = load_oracle(bd), load_ref(bd)
    R.H.refh_sample.argtypes = [ci, vp, ci, ci, ci, vp, ci, ci, ci]
    rng = np.random.default_rng(6100 + bd)
    hi = (1 << bd) - 1
    n = 0
    for kind, name, taps, phases, is_hi in KINDS:
        f = getattr(O, name); f.argtypes = [ci, ci, vp, ci, vp, ci, ci, ci]
        off = 3 if taps == 8 else 1
        for (w, h) in ((4, 4), (8, 8), (16, 8), (8, 16), (32, 32), (64, 64), (4, 16)) if taps == 8 else ((2, 2), (4, 4), (8, 4), (16, 16), (32, 32), (2, 8)):
            S = w + taps + 5
            for ex in (False, True):
                win = (rng.choice([0, hi], (h + taps + 3, S)) if ex else rng.integers(0, hi + 1, (h + taps + 3, S))).astype(pix_dtype(bd))
                src = win[off:, off:]                       # top-left integer sample of the block
                p0 = win.ctypes.data + (off * S + off) * win.itemsize
                step = 1 if w * h <= 64 else 3
                for my in range(0, phases, step):
                    for mx in range(0, phases, step):
                        odt = np.int16 if is_hi else pix_dtype(bd)
                        d1 = np.zeros((h, w), odt); d2 = np.zeros((h, w), odt)
                        f(w, h, p0, S, ptr(d1), w, mx - 64, my + 32)       # integer parts must be ignored (& 15 / & 31)
                        R.H.refh_sample(kind, p0, S, w, h, ptr(d2), w, mx - 64, my + 32)
                        assert (d1 == d2).all(), (bd, name, w, h, mx, my, ex)
                        n += 1
    assert n > 3000


def test_frac_search_batch_matches_oracle():
    """The CPU-baseline leg of tools/ipol_bench.py (refh_frac_search_batch: the reference's
    sample_quarterpel_luma + satd_any_size) against the oracle's interpolation + SATD."""
    bd = 8
    O, R = load_oracle(bd), load_ref(bd)
    f = O.orc_sample_quarterpel_luma; f.argtypes = [ci, ci, vp, ci, vp, ci, ci, ci]
    O.orc_satd_any_size.argtypes = [ci, ci, vp, ci, vp, ci]
    R.H.refh_frac_search_batch.argtypes = [ci, ci, vp, ci, vp, ci, ci, vp, vp, vp, vp, vp, ci, C.c_long, vp]
    rng = np.random.default_rng(6500)
    PW, PH, M = 96, 64, 40
    ref = rng.integers(0, 256, (PH, PW)).astype(np.uint8); cur = rng.integers(0, 256, (PH, PW)).astype(np.uint8)
    padded = np.ascontiguousarray(np.pad(ref, M, mode="edge"))
    offs = np.array([(0, 0), (8, 0), (-8, 8), (4, -4), (-16, 15), (15, -16)], dtype=np.int8)
    k = len(offs)
    for (w, h) in ((8, 8), (16, 16), (32, 16)):
        n = 12
        bx = rng.integers(0, PW - w + 1, n).astype(np.int32); by = rng.integers(0, PH - h + 1, n).astype(np.int32)
        mvx = rng.integers(-20 * 16, 20 * 16, n).astype(np.int32); mvy = rng.integers(-20 * 16, 20 * 16, n).astype(np.int32)
        out = np.zeros(n * k, np.uint32)
        R.H.refh_frac_search_batch(w, h, ptr(cur), PW, ptr(padded), padded.shape[1], M, ptr(bx), ptr(by), ptr(mvx), ptr(mvy),
                                   ptr(offs), k, n, ptr(out))
        for i in range(n):
            orig = np.ascontiguousarray(cur[by[i]:by[i] + h, bx[i]:bx[i] + w])
            for c in range(k):
                cx, cy = int(mvx[i]) + int(offs[c, 0]), int(mvy[i]) + int(offs[c, 1])
                ix, iy = int(bx[i]) + (cx >> 4) + M, int(by[i]) + (cy >> 4) + M
                pred = np.zeros((h, w), np.uint8)
                f(w, h, padded.ctypes.data + iy * padded.shape[1] + ix, padded.shape[1], ptr(pred), w, cx, cy)
                assert out[i * k + c] == O.orc_satd_any_size(w, h, ptr(orig), w, ptr(pred), w), (w, h, i, c)


@pytest.mark.parametrize("bd", [8, 10])
def test_fme_block_filters(bd):
    """uvg_filter_hpel/qpel_blocks_* chained the way search_frac does (refh_fme_filter) against
    orc_fme_filter: every output block is the sample interpolation at one quarter-pel offset."""
    O, R = load_oracle(bd), load_ref(bd)
    O.orc_fme_filter.argtypes = [ci, vp, ci, ci, ci, ci, ci, vp]
    R.H.refh_fme_filter.argtypes = [ci, vp, ci, ci, ci, ci, ci, ci, vp]
    rng = np.random.default_rng(6700 + bd)
    hi = (1 << bd) - 1
    n = 0
    for (w, h) in ((8, 8), (16, 8), (8, 16), (32, 32), (64, 64), (16, 64)):
        S = w + 1 + 16
        for ex in (False, True):
            win = (rng.choice([0, hi], (h + 17, S)) if ex else rng.integers(0, hi + 1, (h + 17, S))).astype(pix_dtype(bd))
            src = win.ctypes.data + (8 * S + 8) * win.itemsize
            for fme_level in (1, 2, 3, 4):
                step = fme_level - 1
                for (ox, oy) in ([(0, 0)] if step < 2 else [(-1, -1), (0, -1), (1, 0), (-1, 1), (0, 0), (1, 1)]):
                    d1 = np.zeros((4, h, w), pix_dtype(bd)); d2 = np.zeros((4, h, w), pix_dtype(bd))
                    O.orc_fme_filter(step, src, S, w, h, ox, oy, ptr(d1))
                    R.H.refh_fme_filter(step, src, S, w, h, fme_level, ox, oy, ptr(d2))
                    assert (d1 == d2).all(), (bd, w, h, fme_level, ox, oy)
                    n += 1
    assert n > 100
