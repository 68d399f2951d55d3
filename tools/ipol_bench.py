"""Fractional motion search scoring of one 3840x2160 8-bit frame (next row 8(f)-1): every w x h block,
17 sub-pel candidates around its motion vector (centre + 8 half-pel + 8 quarter-pel neighbours, the
positions search_frac visits, search_inter.c:1085-1250), cost = satd_any_size(original, interpolated).

Three ways, same inputs:
  fused   : uvg_b200_frac_search_dev - the interpolated blocks never exist in HBM
  2-step  : uvg_b200_ipol_dev -> blocks in HBM -> uvg_b200_distortion_dev (SATD)
  cpu     : the unmodified reference (uvg_sample_quarterpel_luma + satd_any_size, generic C) on a bounded
            sample of the blocks, all host threads the cgroup allows (refh_frac_search_batch)
and checks fused == 2-step over the whole frame and fused == cpu on the sample.

  python tools/ipol_bench.py [--iters 10] [--no-cpu]
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import bench
import uvg266_b200 as U

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--no-cpu", action="store_true")
ap.add_argument("--cpu-seconds", type=float, default=3.0)
args = ap.parse_args()

U.init(0)
L = U.lib()
dev = torch.device("cuda", 0)
W, H = bench.FRAME_W, bench.CTU_ROWS * 64
ref = bench.synth_plane_rows([(0, r) for r in range(bench.CTU_ROWS)])
cur = bench.synth_plane_rows([(1, r) for r in range(bench.CTU_ROWS)])
stream = torch.cuda.current_stream().cuda_stream
d_ref, d_cur = torch.from_numpy(ref).to(dev), torch.from_numpy(cur).to(dev)
offs = [(0, 0)] + [(dx, dy) for dy in (-8, 0, 8) for dx in (-8, 0, 8) if (dx, dy) != (0, 0)] + \
       [(dx, dy) for dy in (-4, 0, 4) for dx in (-4, 0, 4) if (dx, dy) != (0, 0)]
offsets = np.array(offs, dtype=np.int8)
k = len(offs)
rng = np.random.default_rng(7)


def timed(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


res = {"frame": "%dx%d u8" % (W, H), "candidates_per_block": k, "levels": []}
for (w, h) in ((16, 16), (32, 32), (8, 8)):
    ys, xs = np.meshgrid(np.arange(0, H, h), np.arange(0, W, w), indexing="ij")
    bx = xs.ravel().astype(np.int32); by = ys.ravel().astype(np.int32); n = bx.size
    mvx = (rng.integers(-8, 9, n) * 16).astype(np.int32); mvy = (rng.integers(-8, 9, n) * 16).astype(np.int32)   # integer-pel start
    d = [torch.from_numpy(a).to(dev) for a in (bx, by, mvx, mvy)]
    out = torch.zeros(n * k, dtype=torch.int32, device=dev)
    out2 = torch.zeros(n * k, dtype=torch.int32, device=dev)
    preds = torch.empty(n * k * w * h, dtype=torch.uint8, device=dev)
    # 2-step needs one (block, candidate) job per interpolated block
    jb = [torch.from_numpy(np.repeat(a, k)).to(dev) for a in (bx, by)]
    jm = [torch.from_numpy((np.repeat(m, k) + np.tile(offsets[:, i].astype(np.int32), n)).astype(np.int32)).to(dev) for i, m in enumerate((mvx, mvy))]
    A = U.Blocks.plane(d_cur.data_ptr(), W, W, w, h)
    B = U.Blocks.dense(preds.data_ptr(), w, h)

    def fused():
        U.check(L.uvg_b200_frac_search_dev(8, w, h, d_cur.data_ptr(), d_ref.data_ptr(), W, W, H, d[0].data_ptr(), d[1].data_ptr(),
                                           d[2].data_ptr(), d[3].data_ptr(), offsets.ctypes.data, k, out.data_ptr(), n, stream))

    def interp():
        U.check(L.uvg_b200_ipol_dev(8, 0, 0, d_ref.data_ptr(), W, W, H, w, h, jb[0].data_ptr(), jb[1].data_ptr(), jm[0].data_ptr(),
                                    jm[1].data_ptr(), preds.data_ptr(), n * k, stream))

    def score():
        U.check(L.uvg_b200_distortion_dev(U.SATD, 8, w, h, C.byref(A), C.byref(B), out2.data_ptr(), n, k, stream))

    t_f, t_i, t_s = timed(fused, args.iters), timed(interp, args.iters), timed(score, args.iters)
    same = bool(torch.equal(out, out2))
    alg = n * (w * h + (w + 9) * (h + 9) + 4 * k)
    lvl = {"w": w, "h": h, "blocks": int(n), "fused_ms": round(t_f, 4), "interpolate_ms": round(t_i, 4), "score_ms": round(t_s, 4),
           "fused_equals_two_step": same, "candidates_per_s": round(n * k / t_f * 1e3),
           "fused_alg_MB": round(alg / 1e6, 2), "two_step_alg_MB": round((alg + 2 * n * k * w * h) / 1e6, 2)}
    if not args.no_cpu:
        from _libs import load_ref
        R = load_ref(8)
        i32p = C.c_void_p
        R.H.refh_frac_search_batch.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, i32p, i32p, i32p, i32p,
                                               C.c_void_p, C.c_int, C.c_long, C.c_void_p]
        M = 160
        padded = np.ascontiguousarray(np.pad(ref, M, mode="edge"))
        ns = min(n, 40000 * 256 // (w * h))
        sel = np.linspace(0, n - 1, ns).astype(np.int64)
        sb = [np.ascontiguousarray(a[sel]) for a in (bx, by, mvx, mvy)]
        costs = np.zeros(ns * k, np.uint32)
        nt = bench.host_cores()
        bounds = np.linspace(0, ns, nt + 1).astype(int)

        def work(t):
            lo, hi = bounds[t], bounds[t + 1]
            if hi > lo:
                R.H.refh_frac_search_batch(w, h, cur.ctypes.data, W, padded.ctypes.data, padded.shape[1], M, sb[0][lo:].ctypes.data,
                                           sb[1][lo:].ctypes.data, sb[2][lo:].ctypes.data, sb[3][lo:].ctypes.data, offsets.ctypes.data, k,
                                           int(hi - lo), costs[lo * k:].ctypes.data)
        reps, dt = 0, 0.0
        while dt < args.cpu_seconds:
            t0 = time.perf_counter()
            th = [threading.Thread(target=work, args=(t,)) for t in range(nt)]
            [t.start() for t in th]; [t.join() for t in th]
            dt += time.perf_counter() - t0; reps += 1
        dt /= reps
        g = out.cpu().numpy().astype(np.uint32).reshape(n, k)[sel].ravel()
        lvl.update({"cpu_threads": nt, "cpu_sample_blocks": int(ns), "cpu_candidates_per_s": round(ns * k / dt),
                    "cpu_sample_s": round(dt, 3), "gpu_equals_reference_on_sample": bool((g == costs).all()),
                    "speedup_vs_cpu": round((n * k / t_f * 1e3) / (ns * k / dt), 1)})
    res["levels"].append(lvl)
print(json.dumps(res))
