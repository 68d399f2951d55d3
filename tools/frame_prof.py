"""Runs the device-resident frame pass a few times (for ncu launch lists / full captures of its kernels).
  python tools/frame_prof.py [--iters 3]"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import bench
import uvg266_b200 as U

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
U.init(0)
L = U.lib()
H = bench.CTU_ROWS * 64
plane = bench.synth_plane_rows([(0, r) for r in range(bench.CTU_ROWS)])
cfg = U.FrameCfg.make(bench.FRAME_W, H, bench.FRAME_LEVELS, bench.TUS, bench.FRAME_NBEST, U.QuantCfg.make(bench.QP, 8, 1, 0, 0),
                      bench.COST_WEIGHTS)
f = L.uvg_b200_frame_create(C.byref(cfg))
res = U.FrameResult()
U.check(L.uvg_b200_frame_submit(f, plane.ctypes.data, bench.FRAME_W, 0, 0, 0))
U.check(L.uvg_b200_frame_wait(f, C.byref(res)))
ms = []
for _ in range(a.iters):
    U.check(L.uvg_b200_frame_submit(f, None, bench.FRAME_W, 0, 0, U.FRAME_NO_H2D | U.FRAME_NO_D2H))
    U.check(L.uvg_b200_frame_wait(f, C.byref(res)))
    ms.append(res.device_ms)
print("device ms per frame:", ms)
L.uvg_b200_frame_destroy(f)
