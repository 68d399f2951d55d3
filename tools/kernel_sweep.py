"""Kernel benchmark of SURVEY.md 8(d)(i): independent block pairs / TUs, data >> L2 (no reuse), every
size and both bit depths, algorithmic GB/s against the measured HBM peak.

  distortion : SAD / SATD / SSD, n x n pairs (n = 4..64), 8-bit and 10-bit, ~1 GiB read per launch
  transform  : forward and inverse DCT-2, n = 4..32, 2^20 TUs (4*n*n bytes per TU)
  quant      : quant and dequant, n = 4..32, 2^20 TUs, QP 22 / 27 / 32 / 37 (timing is QP independent: one row)

  python tools/kernel_sweep.py [--only transform] > profiles/r2_kernel_sweep.json
  UVG266_B200_TRANSFORM=fast python tools/kernel_sweep.py --only transform      # the packed-ALU transform kernel
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import uvg266_b200 as U

U.init(0)
L = U.lib()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except (OSError, KeyError):
    PEAK = 6650.0


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


ONLY = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else ""      # e.g. --only transform
res = {"hbm_peak_gbs": PEAK, "distortion": [], "transform": [], "quant": [], "transform_kernel": os.environ.get("UVG266_B200_TRANSFORM", "imma (16x16, 32x32) / fast")}
g = torch.Generator(device=dev); g.manual_seed(1234)
for bd in (() if ONLY and ONLY != "distortion" else (8, 10)):
    s = 1 if bd == 8 else 2
    total = 1 << 29                      # bytes per operand: 512 MiB (1 GiB read per launch)
    a = torch.randint(0, 1 << bd, (total // s,), dtype=torch.int16 if bd != 8 else torch.uint8, device=dev, generator=g)
    b = torch.randint(0, 1 << bd, (total // s,), dtype=torch.int16 if bd != 8 else torch.uint8, device=dev, generator=g)
    for n in (4, 8, 16, 32, 64):
        nb = total // (n * n * s)
        A, B = U.Blocks.dense(a.data_ptr(), n, n), U.Blocks.dense(b.data_ptr(), n, n)
        out = torch.zeros(nb, dtype=torch.int32, device=dev)
        for name, metric in (("SAD", U.SAD_NXN), ("SATD", U.SATD_4X4 if n == 4 else U.SATD), ("SSD", U.SSD)):
            ms = timed(lambda: U.check(L.uvg_b200_distortion_dev(metric, bd, n, n, C.byref(A), C.byref(B), out.data_ptr(), nb, 1, stream)))
            gb = nb * (2 * n * n * s + 4) / ms / 1e6
            res["distortion"].append({"metric": name, "n": n, "bitdepth": bd, "pairs": nb, "ms": round(ms, 4), "gbps": round(gb, 1),
                                      "frac": round(gb / PEAK, 3)})
    del a, b
NT = 1 << 20
cfg = U.QuantCfg.make(27, 8, 1, 0, 0)
for n in (() if ONLY and ONLY not in ("transform", "quant") else (4, 8, 16, 32)):
    src = torch.randint(-255, 256, (NT * n * n,), dtype=torch.int16, device=dev, generator=g)
    dst = torch.zeros_like(src)
    for inv in (0, 1):
        ms = timed(lambda: U.check(L.uvg_b200_transform_dev(inv, 8, n, n, 0, 0, 0, 0, src.data_ptr(), dst.data_ptr(), NT, stream)))
        gb = NT * 4 * n * n / ms / 1e6
        res["transform"].append({"op": "idct" if inv else "dct", "n": n, "tus": NT, "ms": round(ms, 4), "gbps": round(gb, 1), "frac": round(gb / PEAK, 3),
                                 "gmac_per_s": round(NT * 2 * n * n * n / ms / 1e6, 1)})
    ms = timed(lambda: U.check(L.uvg_b200_quant_dev(C.byref(cfg), src.data_ptr(), dst.data_ptr(), n, n, 0, 0, 0, NT, stream)))
    res["quant"].append({"op": "quant", "n": n, "tus": NT, "ms": round(ms, 4), "gbps": round(NT * 4 * n * n / ms / 1e6, 1),
                         "frac": round(NT * 4 * n * n / ms / 1e6 / PEAK, 3)})
    ms = timed(lambda: U.check(L.uvg_b200_dequant_dev(C.byref(cfg), dst.data_ptr(), src.data_ptr(), n, n, 0, 0, NT, stream)))
    res["quant"].append({"op": "dequant", "n": n, "tus": NT, "ms": round(ms, 4), "gbps": round(NT * 4 * n * n / ms / 1e6, 1),
                         "frac": round(NT * 4 * n * n / ms / 1e6 / PEAK, 3)})
    del src, dst
print(json.dumps(res))
