"""GPU parity, next row 8(f)-1: MC sample interpolation (drop-in leaves, batched entry point) and the fused
fractional-search scoring kernel against the oracle (pinned to the reference by tests/test_oracle_ipol_vs_ref.py)."""
import ctypes as C

import numpy as np
import pytest
import torch

import uvg266_b200 as U
from _libs import load_oracle, ptr, pix_dtype

pytestmark = pytest.mark.gpu
ci, vp = C.c_int, C.c_void_p
KINDS = [("sample_quarterpel_luma", "orc_sample_quarterpel_luma", 8, 16, False),
         ("sample_quarterpel_luma_hi", "orc_sample_quarterpel_luma_hi", 8, 16, True),
         ("sample_octpel_chroma", "orc_sample_octpel_chroma", 4, 32, False),
         ("sample_octpel_chroma_hi", "orc_sample_octpel_chroma_hi", 4, 32, True)]


@pytest.fixture(scope="module", autouse=True)
def _init():
    U.init(0)
    U.lib().uvg_b200_set_dropin_bitdepth(10)


@pytest.mark.parametrize("bd", [8, 10])
def test_sample_leaf_dropins(bd):
    O, L = load_oracle(bd), U.lib()
    rng = np.random.default_rng(6200 + bd)
    hi = (1 << bd) - 1
    for name, oname, taps, phases, is_hi in KINDS:
        f = getattr(O, oname); f.argtypes = [ci, ci, vp, ci, vp, ci, ci, ci]
        g = U.dropin(name, bd)
        off = 3 if taps == 8 else 1
        for (w, h) in ((4, 4), (8, 8), (16, 8), (32, 32), (64, 64), (8, 64)) if taps == 8 else ((2, 2), (4, 4), (8, 4), (32, 32), (2, 16)):
            S = w + taps + 9
            win = (rng.choice([0, hi], (h + taps + 3, S)) if w == 8 else rng.integers(0, hi + 1, (h + taps + 3, S))).astype(pix_dtype(bd))
            p0 = win.ctypes.data + (off * S + off) * win.itemsize
            for (mx, my) in [(0, 0), (1, 0), (0, 5), (phases // 2, phases // 2), (phases - 1, 3), (7, phases - 1), (-3, 37)]:
                odt = np.int16 if is_hi else pix_dtype(bd)
                ds = w + 5                                   # strided destination like the encoder's LCU buffers
                d1 = np.zeros((h, ds), odt); d2 = np.zeros((h, ds), odt)
                f(w, h, p0, S, ptr(d1), ds, mx, my)
                g(p0, S, w, h, ptr(d2), ds, mx, my)
                assert (d1 == d2).all(), (bd, name, w, h, mx, my)


def replicate(plane, m):
    return np.pad(plane, m, mode="edge")


@pytest.mark.parametrize("bd", [8, 10])
def test_batched_ipol_plane_with_edges(bd):
    """Blocks read straight from a picture plane, MVs pointing partly outside: border replication."""
    O, L = load_oracle(bd), U.lib()
    rng = np.random.default_rng(6300 + bd)
    hi = (1 << bd) - 1
    PW, PH, M = 96, 64, 80
    plane = rng.integers(0, hi + 1, (PH, PW)).astype(pix_dtype(bd))
    padded = np.ascontiguousarray(replicate(plane, M))
    tdt = torch.uint8 if bd == 8 else torch.int16
    d_plane = torch.from_numpy(plane.view(np.uint8 if bd == 8 else np.int16)).cuda()
    stream = torch.cuda.current_stream().cuda_stream
    for chroma, oname_p, oname_h, taps in ((0, "orc_sample_quarterpel_luma", "orc_sample_quarterpel_luma_hi", 8),
                                           (1, "orc_sample_octpel_chroma", "orc_sample_octpel_chroma_hi", 4)):
        frac = 4 if taps == 8 else 5
        for (w, h) in ((16, 16), (8, 32)):
            n = 40
            bx = rng.integers(0, PW - w + 1, n).astype(np.int32); by = rng.integers(0, PH - h + 1, n).astype(np.int32)
            mvx = rng.integers(-40 << frac, 40 << frac, n).astype(np.int32); mvy = rng.integers(-40 << frac, 40 << frac, n).astype(np.int32)
            mvx[:4] = [-(PW + 20) << frac, (PW + 20) << frac, 3, 0]; mvy[:4] = [5, -(PH + 30) << frac, (PH + 30) << frac, 0]
            mvx = np.clip(mvx, -(M - 12 - PW // 2) << frac, (M - 12) << frac); mvy = np.clip(mvy, -(M - 12) << frac, (M - 12) << frac)
            d = [torch.from_numpy(a).cuda() for a in (bx, by, mvx, mvy)]
            for is_hi, oname in ((0, oname_p), (1, oname_h)):
                f = getattr(O, oname); f.argtypes = [ci, ci, vp, ci, vp, ci, ci, ci]
                out = torch.zeros(n * w * h, dtype=torch.int16 if (is_hi or bd != 8) else torch.uint8, device="cuda")
                U.check(L.uvg_b200_ipol_dev(bd, chroma, is_hi, d_plane.data_ptr(), PW, PW, PH, w, h, d[0].data_ptr(), d[1].data_ptr(),
                                            d[2].data_ptr(), d[3].data_ptr(), out.data_ptr(), n, stream))
                torch.cuda.synchronize()
                got = out.cpu().numpy().reshape(n, h, w)
                for i in range(n):
                    ix, iy = int(bx[i]) + (int(mvx[i]) >> frac) + M, int(by[i]) + (int(mvy[i]) >> frac) + M
                    # clamp the window origin the way coordinate clamping does (far outside: constant border)
                    ix = min(max(ix, 8), padded.shape[1] - w - 8); iy = min(max(iy, 8), padded.shape[0] - h - 8)
                    p0 = padded.ctypes.data + (iy * padded.shape[1] + ix) * padded.itemsize
                    e = np.zeros((h, w), np.int16 if is_hi else pix_dtype(bd))
                    f(w, h, p0, padded.shape[1], ptr(e), w, int(mvx[i]), int(mvy[i]))
                    g = got[i].astype(e.dtype) if not is_hi else got[i]
                    if bd != 8 and not is_hi:
                        g = got[i].view(np.uint16)
                    assert (g == e).all(), (bd, chroma, is_hi, w, h, i, int(mvx[i]), int(mvy[i]))


def test_fused_frac_search():
    bd = 8
    O, L = load_oracle(bd), U.lib()
    f = O.orc_sample_quarterpel_luma; f.argtypes = [ci, ci, vp, ci, vp, ci, ci, ci]
    O.orc_satd_any_size.argtypes = [ci, ci, vp, ci, vp, ci]
    rng = np.random.default_rng(6400)
    PW, PH, M = 128, 96, 32
    base = np.clip(128 + np.cumsum(rng.integers(-5, 6, (PH, PW)), axis=1), 0, 255).astype(np.uint8)
    ref = base.copy(); cur = np.roll(base, (1, 2), (0, 1))
    cur = np.clip(cur.astype(int) + rng.integers(-3, 4, cur.shape), 0, 255).astype(np.uint8)
    padded = np.ascontiguousarray(replicate(ref, M))
    stream = torch.cuda.current_stream().cuda_stream
    d_ref, d_cur = torch.from_numpy(ref).cuda(), torch.from_numpy(cur).cuda()
    # centre + 8 half-pel + 8 quarter-pel neighbours (search_inter.c:1100-1250), in 1/16 samples
    offs = [(0, 0)] + [(dx, dy) for dy in (-8, 0, 8) for dx in (-8, 0, 8) if (dx, dy) != (0, 0)] + \
           [(dx, dy) for dy in (-4, 0, 4) for dx in (-4, 0, 4) if (dx, dy) != (0, 0)] + [(-16, 15), (15, -16)]
    offsets = np.array(offs, dtype=np.int8)
    k = len(offs)
    for (w, h) in ((8, 8), (16, 16), (32, 16), (64, 64)):
        ys, xs = np.meshgrid(np.arange(0, PH - h + 1, h), np.arange(0, PW - w + 1, w), indexing="ij")
        bx = xs.ravel().astype(np.int32); by = ys.ravel().astype(np.int32); n = bx.size
        mvx = (rng.integers(-6, 7, n) * 16 + rng.integers(0, 16, n)).astype(np.int32)
        mvy = (rng.integers(-6, 7, n) * 16 + rng.integers(0, 16, n)).astype(np.int32)
        d = [torch.from_numpy(a).cuda() for a in (bx, by, mvx, mvy)]
        out = torch.zeros(n * k, dtype=torch.int32, device="cuda")
        U.check(L.uvg_b200_frac_search_dev(8, w, h, d_cur.data_ptr(), d_ref.data_ptr(), PW, PW, PH, d[0].data_ptr(), d[1].data_ptr(),
                                           d[2].data_ptr(), d[3].data_ptr(), ptr(offsets), k, out.data_ptr(), n, stream))
        torch.cuda.synchronize()
        got = out.cpu().numpy().reshape(n, k)
        for i in range(0, n, max(1, n // 10)):
            orig = np.ascontiguousarray(cur[by[i]:by[i] + h, bx[i]:bx[i] + w])
            for c, (dx, dy) in enumerate(offs):
                cx, cy = int(mvx[i]) + dx, int(mvy[i]) + dy
                ix, iy = int(bx[i]) + (cx >> 4) + M, int(by[i]) + (cy >> 4) + M
                p0 = padded.ctypes.data + iy * padded.shape[1] + ix
                pred = np.zeros((h, w), np.uint8)
                f(w, h, p0, padded.shape[1], ptr(pred), w, cx, cy)
                e = O.orc_satd_any_size(w, h, ptr(orig), w, ptr(pred), w)
                assert got[i, c] == e, (w, h, i, c, dx, dy)


@pytest.mark.parametrize("bd", [8, 10])
def test_fme_filter_dropins(bd):
    """The four FME block filters (filter_hpel/qpel_blocks_*) as drop-ins, output layout of the reference
    (filtered[4][LCU_LUMA_SIZE], LCU_WIDTH stride), against orc_fme_filter."""
    O, L = load_oracle(bd), U.lib()
    O.orc_fme_filter.argtypes = [ci, vp, ci, ci, ci, ci, ci, vp]
    g = U.dropin("fme_filter", bd)
    rng = np.random.default_rng(6800 + bd)
    hi = (1 << bd) - 1
    LW, LSZ = 64, 64 * 64
    for (w, h) in ((8, 8), (16, 8), (32, 32), (64, 64), (8, 32)):
        S = w + 1 + 16
        win = rng.integers(0, hi + 1, (h + 17, S)).astype(pix_dtype(bd))
        src = win.ctypes.data + (8 * S + 8) * win.itemsize
        for step in range(4):
            for (ox, oy) in ([(0, 0)] if step < 2 else [(-1, -1), (1, 0), (0, 1), (0, 0)]):
                e = np.zeros((4, h, w), pix_dtype(bd))
                O.orc_fme_filter(step, src, S, w, h, ox, oy, ptr(e))
                out = np.zeros((4, LSZ), pix_dtype(bd))
                g(step, src, S, w, h, ptr(out), LW, LSZ, ox, oy)
                got = out.reshape(4, 64, 64)[:, :h, :w]
                assert (got == e).all(), (bd, w, h, step, ox, oy)
