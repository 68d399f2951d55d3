"""Rough intra search of one 3840x2160 8-bit frame (next row 8(f)-2): every n x n block of the
frame, k candidate luma modes each, cost = min(SATD, 2*SAD) (search_intra.c:1055-1062, :149-163).

Three ways, same inputs (reference rows taken from the neighbouring pixels of the same plane):
  fused   : uvg_b200_intra_rough_dev - the predictions never exist in HBM
  2-step  : uvg_b200_intra_predict_dev -> predictions in HBM -> uvg_b200_rough_cost_dev
  cpu     : the unmodified reference (uvg_intra_predict + satd/sad duals, generic C) on a bounded
            sample of the blocks, all host threads the cgroup allows (refh_intra_rough_batch)
and checks fused == 2-step over the whole frame and fused == cpu on the sample.

  python tools/intra_bench.py [--iters 20] [--no-cpu]
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
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--no-cpu", action="store_true")
ap.add_argument("--cpu-blocks", type=int, default=400000)
ap.add_argument("--cpu-seconds", type=float, default=3.0)
args = ap.parse_args()

U.init(0)
L = U.lib()
dev = torch.device("cuda", 0)
W, H = bench.FRAME_W, bench.CTU_ROWS * 64
plane = bench.synth_plane_rows([(0, r) for r in range(bench.CTU_ROWS)])       # [H][W] u8
stream = torch.cuda.current_stream().cuda_stream
LEVELS = [(8, list(range(0, 67, 3))), (16, list(range(0, 67, 3)) + [1]), (32, list(range(0, 67, 4)))]


def build_refs(n):
    """[nb][2][2n+4] u8: the row above / the column left of each block (clamped at the frame edge)."""
    gp = 2 * n + 4
    ys, xs = np.meshgrid(np.arange(0, H, n), np.arange(0, W, n), indexing="ij")
    ys, xs = ys.ravel(), xs.ravel()
    i = np.arange(gp)
    ty = np.clip(ys - 1, 0, H - 1)[:, None]; tx = np.clip(xs[:, None] - 1 + i[None, :], 0, W - 1)
    top = plane[ty, tx]
    ly = np.clip(ys[:, None] - 1 + i[None, :], 0, H - 1); lx = np.clip(xs - 1, 0, W - 1)[:, None]
    left = plane[ly, lx]
    return np.ascontiguousarray(np.stack([top, left], axis=1))


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


d_plane = torch.from_numpy(plane).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {"frame": "%dx%d u8" % (W, H), "levels": []}
peak = bench.hbm_peak()[0] if hasattr(bench, "hbm_peak") else None
for n, modes in LEVELS:
    k = len(modes)
    refs = build_refs(n)
    nb = refs.shape[0]
    d_refs = torch.from_numpy(refs).to(dev)
    m8 = np.array(modes, dtype=np.int8)
    desc = U.Blocks.plane(d_plane.data_ptr(), W, W, n, n)
    d_cost = torch.zeros(nb * k, dtype=torch.int32, device=dev)
    d_cost2 = torch.zeros(nb * k, dtype=torch.int32, device=dev)
    d_pred = torch.empty(nb * k * n * n, dtype=torch.uint8, device=dev)
    pdesc = U.Blocks.dense(d_pred.data_ptr(), n, n)

    def fused():
        U.check(L.uvg_b200_intra_rough_dev(8, n, C.byref(desc), d_refs.data_ptr(), m8.ctypes.data, k,
                                           d_cost.data_ptr(), nb, stream))

    def predict():
        U.check(L.uvg_b200_intra_predict_dev(8, n, d_refs.data_ptr(), m8.ctypes.data, k, d_pred.data_ptr(), nb, stream))

    def score():
        U.check(L.uvg_b200_rough_cost_dev(8, n, C.byref(desc), C.byref(pdesc), d_cost2.data_ptr(), nb, k, stream))

    t_f = timed(fused, args.iters); t_p = timed(predict, args.iters); t_s = timed(score, args.iters)
    same = bool(torch.equal(d_cost, d_cost2))
    alg = nb * (n * n + 2 * (2 * n + 4) + 4 * k)          # orig + refs in, costs out
    lvl = {"n": n, "modes": k, "blocks": nb, "fused_ms": round(t_f, 4), "predict_ms": round(t_p, 4),
           "score_ms": round(t_s, 4), "fused_equals_two_step": same,
           "candidates_per_s": round(nb * k / t_f * 1e3), "fused_alg_MB": round(alg / 1e6, 2),
           "two_step_alg_MB": round((alg + 2 * nb * k * n * n) / 1e6, 2),
           "pred_pixels_per_s": round(nb * k * n * n / t_f * 1e3)}
    if not args.no_cpu:
        from _libs import load_ref, ptr
        R = load_ref(8)
        R.H.refh_intra_rough_batch.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
        ns = min(nb, args.cpu_blocks * 64 // (n * n))
        sel = np.linspace(0, nb - 1, ns).astype(np.int64)
        ys, xs = (sel // (W // n)) * n, (sel % (W // n)) * n
        orig = np.ascontiguousarray(np.stack([plane[y:y + n, x:x + n].ravel() for y, x in zip(ys, xs)]))
        srefs = np.ascontiguousarray(refs[sel])
        costs = np.zeros(ns * k, np.uint32)
        nt = bench.host_cores()
        bounds = np.linspace(0, ns, nt + 1).astype(int)

        def work(t):
            lo, hi = bounds[t], bounds[t + 1]
            if hi > lo:
                R.H.refh_intra_rough_batch(n, m8.ctypes.data, k, srefs[lo:hi].ctypes.data, orig[lo:hi].ctypes.data,
                                           int(hi - lo), costs[lo * k:].ctypes.data)
        reps, dt = 0, 0.0
        while dt < args.cpu_seconds:                      # repeat the sample until the leg ran long enough
            t0 = time.perf_counter()
            th = [threading.Thread(target=work, args=(t,)) for t in range(nt)]
            [t.start() for t in th]; [t.join() for t in th]
            dt += time.perf_counter() - t0; reps += 1
        dt /= reps
        g = d_cost.cpu().numpy().astype(np.uint32).reshape(nb, k)[sel].ravel()
        lvl.update({"cpu_threads": nt, "cpu_sample_blocks": int(ns), "cpu_candidates_per_s": round(ns * k / dt),
                    "cpu_sample_s": round(dt, 3), "cpu_reps": reps, "gpu_equals_reference_on_sample": bool((g == costs).all()),
                    "speedup_vs_cpu": round((nb * k / t_f * 1e3) / (ns * k / dt), 1)})
    res["levels"].append(lvl)
tot = sum(l["fused_ms"] for l in res["levels"])
res["frame_ms_fused"] = round(tot, 4)
res["frame_ms_two_step"] = round(sum(l["predict_ms"] + l["score_ms"] for l in res["levels"]), 4)
res["launch_count"] = int(L.uvg_b200_launch_count())
print(json.dumps(res))
