"""End-to-end variant of the bench step with the rough-search predictions generated on the device
(third row): per 4K frame
  rough search : every 8x8 block x 23 modes and every 16x16 block x 24 modes, cost = min(SATD, 2 SAD);
                 inputs = the source plane + the blocks' reference rows (instead of 425 MB of predictions)
  TUs          : the bench's 240 x 4x4 + 120 x 8x8 + 24 x 16x16 + 5 x 32x32 per CTU through quantize_residual
GPU arm : uvg_b200_intra_rough_host + uvg_b200_quantize_residual_host from pinned host buffers, H2D and D2H
          inside the timed region, T host threads (same structure as bench.py's e2e).
CPU arm : the unmodified reference on the host threads the cgroup allows: uvg_intra_predict + satd/sad duals
          for every candidate (refh_intra_rough_batch) and the TU chain (refh_ctu_pass without rough levels).

  python tools/e2e_predict_on_device.py [--steps 6] [--threads 4]
"""
import argparse
import concurrent.futures
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


def run(steps=6, threads=4, cpu_arm=True):
    U.init(0)
    L = U.lib()
    W, H = bench.FRAME_W, bench.CTU_ROWS * 64
    n_ctu = bench.CTU_COLS * bench.CTU_ROWS
    plane = bench.synth_plane_rows([(0, r) for r in range(bench.CTU_ROWS)])
    T = threads
    cfg = U.QuantCfg.make(bench.QP, 8, 1, 0, 0)
    LEVELS = [(8, list(range(0, 67, 3))), (16, list(range(0, 67, 3)) + [1])]


    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a[:1]).dtype, pin_memory=True)
        t.copy_(torch.from_numpy(a))
        return t


    def build_refs(n):
        gp = 2 * n + 4
        ys, xs = np.meshgrid(np.arange(0, H, n), np.arange(0, W, n), indexing="ij")
        ys, xs = ys.ravel(), xs.ravel()
        i = np.arange(gp)
        top = plane[np.clip(ys - 1, 0, H - 1)[:, None], np.clip(xs[:, None] - 1 + i[None, :], 0, W - 1)]
        left = plane[np.clip(ys[:, None] - 1 + i[None, :], 0, H - 1), np.clip(xs - 1, 0, W - 1)[:, None]]
        return np.ascontiguousarray(np.stack([top, left], axis=1))


    h_plane = pinned(plane)
    rough = []
    for n, modes in LEVELS:
        refs = build_refs(n)
        nb = refs.shape[0]
        rough.append(dict(n=n, k=len(modes), nb=nb, m8=np.array(modes, dtype=np.int8), refs_np=refs, refs=pinned(refs),
                          out=torch.empty(nb * len(modes), dtype=torch.int32, pin_memory=True), bpr=W // n, brows=H // n))
    rng = np.random.default_rng(99)
    tus = []
    for (n, cnt) in bench.TUS:
        r = bench.gather_blocks(plane, n, cnt)
        p = bench.noisy(r, rng, 24)
        nt = n_ctu * cnt
        tus.append(dict(n=n, nt=nt, r_np=r, p_np=p, r=pinned(r), p=pinned(p), rec=torch.empty(r.size, dtype=torch.uint8, pin_memory=True),
                        coeff=torch.empty(r.size, dtype=torch.int16, pin_memory=True), has=torch.empty(nt, dtype=torch.uint8, pin_memory=True),
                        ssd=torch.empty(nt, dtype=torch.int32, pin_memory=True), cost=torch.empty(nt, dtype=torch.int32, pin_memory=True)))
    h2d = plane.size * len(rough) + sum(x["refs_np"].size for x in rough) + sum(2 * x["r_np"].size for x in tus)
    d2h = sum(x["nb"] * x["k"] * 4 for x in rough) + sum(3 * x["r_np"].size + x["nt"] * 9 for x in tus)


    def work(w):
        # (taking the TU slice first on odd threads, as bench.py's e2e does, measured 428 k against 445 k here:
        # this step is download-bound in both halves, so the order stays rough -> TUs)
        do_rough(w); do_tus(w)

    def do_rough(w):
        for x in rough:
            n, k = x["n"], x["k"]
            r0, r1 = x["brows"] * w // T, x["brows"] * (w + 1) // T              # block rows of this thread
            b0, nbk = r0 * x["bpr"], (r1 - r0) * x["bpr"]
            desc = U.Blocks.plane(h_plane.data_ptr() + r0 * n * W, W, W, n, n)
            U.check(L.uvg_b200_intra_rough_host(8, n, C.byref(desc), (r1 - r0) * n * W, x["refs"].data_ptr() + b0 * 2 * (2 * n + 4),
                                                x["m8"].ctypes.data, k, x["out"].data_ptr() + b0 * k * 4, nbk))
    def do_tus(w):
        for x in tus:
            n = x["n"]; nn = n * n
            lo, hi = x["nt"] * w // T, x["nt"] * (w + 1) // T
            R = U.Blocks.dense(x["r"].data_ptr() + lo * nn, n, n); P = U.Blocks.dense(x["p"].data_ptr() + lo * nn, n, n)
            O = U.Blocks.dense(x["rec"].data_ptr() + lo * nn, n, n)
            U.check(L.uvg_b200_quantize_residual_host(C.byref(cfg), 1, 0, n, n, 0, 0, 0, C.byref(R), (hi - lo) * nn, C.byref(P), (hi - lo) * nn,
                                                      C.byref(O), (hi - lo) * nn, x["coeff"].data_ptr() + lo * nn * 2, x["has"].data_ptr() + lo,
                                                      x["ssd"].data_ptr() + lo * 4, x["cost"].data_ptr() + lo * 4, bench.COST_WEIGHTS, hi - lo))


    pool = concurrent.futures.ThreadPoolExecutor(T)
    for _ in range(2):
        list(pool.map(work, range(T)))
    t0 = time.perf_counter()
    for _ in range(steps):
        list(pool.map(work, range(T)))
    gpu_s = (time.perf_counter() - t0) / steps
    res = {"frame": "%dx%d u8, %d CTUs" % (W, H, n_ctu), "gpu": {"ctus_per_s": round(n_ctu / gpu_s), "ms_per_frame": round(gpu_s * 1e3, 3),
                                                              "h2d_bytes": int(h2d), "d2h_bytes": int(d2h), "host_threads": T}}

    if not cpu_arm:
        return res
    # ---- CPU arm: the unmodified reference -------------------------------------------------------------
    from _libs import load_ref
    arm = bench.CpuArm(n_ctu)
    R = arm.R
    R.H.refh_intra_rough_batch.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
    nt = bench.host_cores()
    cpu_rough = 0.0
    for x in rough:
        n, k, nb = x["n"], x["k"], x["nb"]
        orig = np.ascontiguousarray(plane.reshape(H // n, n, W // n, n).transpose(0, 2, 1, 3)).reshape(nb, n * n)
        costs = np.zeros(nb * k, np.uint32)
        bounds = np.linspace(0, nb, nt + 1).astype(int)

        def wk(t, x=x, orig=orig, costs=costs, bounds=bounds, n=n, k=k):
            lo, hi = bounds[t], bounds[t + 1]
            R.H.refh_intra_rough_batch(n, x["m8"].ctypes.data, k, x["refs_np"][lo:].ctypes.data, orig[lo:].ctypes.data, int(hi - lo),
                                       costs[lo * k:].ctypes.data)
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            th = [threading.Thread(target=wk, args=(t,)) for t in range(nt)]
            [t.start() for t in th]; [t.join() for t in th]
            best = min(best, time.perf_counter() - t0)
        g = x["out"].numpy().astype(np.uint32)
        x["match"] = bool((g == costs).all())
        cpu_rough += best
    (ro, rp, rc), (tr, tp, trec, tco, tss, tcs) = arm.args
    tu_s = min(R.H.refh_ctu_pass(arm.strategy, nt, n_ctu, 0, arm.rough, ro, rp, rc, len(bench.TUS), arm.tu, tr, tp, trec, tco, tss, tcs,
                                 bench.COST_WEIGHTS) for _ in range(3))
    res["cpu_reference"] = {"threads": nt, "rough_with_prediction_s": round(cpu_rough, 4), "tus_s": round(tu_s, 4),
                            "ctus_per_s": round(n_ctu / (cpu_rough + tu_s))}
    res["gpu_costs_equal_reference"] = all(x["match"] for x in rough)
    res["speedup"] = round((n_ctu / gpu_s) / (n_ctu / (cpu_rough + tu_s)), 1)
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true", help="GPU arm only")
    a = ap.parse_args()
    print(json.dumps(run(a.steps, a.threads, not a.no_cpu)))
