"""Host-side timeline of the end-to-end frame pipeline: `--inflight` threads, each looping submit/wait on its own frame
object; prints per picture the submit begin/end, the wait end (ms since start) and the pass's device time, so that idle
gaps and lock-step between the pictures in flight are visible without a system profiler.
  python tools/e2e_timeline.py [--inflight 2] [--steps 6] [--stagger-ms 0]"""
import argparse
import ctypes as C
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import uvg266_b200 as U

ap = argparse.ArgumentParser()
ap.add_argument("--inflight", type=int, default=2)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--stagger-ms", type=float, default=0.0)
ap.add_argument("--flags", type=int, default=0)
a = ap.parse_args()
L = U.lib()
H = bench.CTU_ROWS * bench.CTU
cfg = U.FrameCfg.make(bench.FRAME_W, H, bench.FRAME_LEVELS, bench.TUS, bench.FRAME_NBEST, U.QuantCfg.make(bench.QP, 8, 1, 0, 0), bench.COST_WEIGHTS)
frames, planes = [], []
for i in range(a.inflight):
    f = L.uvg_b200_frame_create(C.byref(cfg)); assert f
    hp = torch.empty((H, bench.FRAME_W), dtype=torch.uint8, pin_memory=True)
    hp.copy_(torch.from_numpy(bench.synth_plane_rows([(i, r) for r in range(bench.CTU_ROWS)])))
    frames.append(f); planes.append(hp)
r = U.FrameResult()
for i in range(a.inflight):
    for _ in range(3):
        U.check(L.uvg_b200_frame_submit(frames[i], planes[i].data_ptr(), bench.FRAME_W, 0, 0, a.flags)); U.check(L.uvg_b200_frame_wait(frames[i], C.byref(r)))
rows = []
t00 = time.perf_counter()


def worker(i):
    rr = U.FrameResult()
    if a.stagger_ms:
        time.sleep(i * a.stagger_ms * 1e-3)
    for s in range(a.steps):
        t0 = time.perf_counter()
        U.check(L.uvg_b200_frame_submit(frames[i], planes[i].data_ptr(), bench.FRAME_W, 0, 0, a.flags))
        t1 = time.perf_counter()
        U.check(L.uvg_b200_frame_wait(frames[i], C.byref(rr)))
        t2 = time.perf_counter()
        rows.append((t0 - t00, i, s, t1 - t0, t2 - t00, rr.device_ms))


th = [threading.Thread(target=worker, args=(i,)) for i in range(a.inflight)]
for t in th: t.start()
for t in th: t.join()
total = time.perf_counter() - t00
print("inflight %d steps %d stagger %.2f ms: %.3f ms per picture" % (a.inflight, a.steps, a.stagger_ms, total * 1e3 / (a.inflight * a.steps)))
print("frame step | submit at | submit took | wait returned | device_ms")
for (t0, i, s, ds, t2, dm) in sorted(rows):
    print("  %d    %2d  | %8.3f | %8.3f | %8.3f | %6.3f" % (i, s, t0 * 1e3, ds * 1e3, t2 * 1e3, dm), flush=False)
