"""Search-shaped benchmark + the one collective of the inter path (8e).

Per frame: every 64x64 CTU of this rank's band of a 3840x2160 8-bit luma plane is searched over
the full (2r+1)^2 integer window around MV 0 in the previous frame (uvg_b200_sad_search_dev:
reg_sad / ver_sad / hor_sad semantics), then the ranks all-gather their bands of the new
"reconstructed" picture (here: the current frame) so that every GPU holds the whole reference for
the next frame (uvg266_b200.refpic.exchange_reference, NCCL over NVLink).

  python tools/search_bench.py [--range 32] [--frames 8]
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/search_bench.py
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
import bench
import uvg266_b200 as U
from uvg266_b200 import refpic

ap = argparse.ArgumentParser()
ap.add_argument("--range", type=int, default=32)
ap.add_argument("--frames", type=int, default=8)
ap.add_argument("--block", type=int, default=64)
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
U.init(local)
L = U.lib()
dev = torch.device("cuda", local)
W, Hpad, n = bench.FRAME_W, bench.CTU_ROWS * 64, args.block
frames = [torch.from_numpy(bench.synth_plane_rows([(f, r) for r in range(bench.CTU_ROWS)])).to(dev) for f in range(2)]
lo, hi = refpic.band_rows(rank, world, bench.CTU_ROWS)
ys, xs = np.meshgrid(np.arange(lo * 64, hi * 64, n), np.arange(0, W, n), indexing="ij")
cx = torch.from_numpy(xs.ravel().astype(np.int32)).to(dev); cy = torch.from_numpy(ys.ravel().astype(np.int32)).to(dev)
zero = torch.zeros_like(cx)
nb = cx.numel(); side = 2 * args.range + 1
out = torch.zeros(nb * side * side, dtype=torch.int32, device=dev)
stream = torch.cuda.current_stream().cuda_stream
ref = frames[0].clone(); cur = frames[1]


def frame():
    U.check(L.uvg_b200_sad_search_dev(8, n, n, cur.data_ptr(), ref.data_ptr(), W, Hpad, W, cx.data_ptr(), cy.data_ptr(),
                                      zero.data_ptr(), zero.data_ptr(), args.range, out.data_ptr(), nb, stream))
    # this rank's band of the new reference, then one collective
    ref[lo * 64: hi * 64] = cur[lo * 64: hi * 64]
    if world > 1:
        refpic.exchange_reference(ref, rank, world)


for _ in range(3):
    frame()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.frames):
    frame()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / args.frames], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ms = float(ms.item())
# search only (no exchange), for the kernel figure
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(args.frames):
    U.check(L.uvg_b200_sad_search_dev(8, n, n, cur.data_ptr(), ref.data_ptr(), W, Hpad, W, cx.data_ptr(), cy.data_ptr(),
                                      zero.data_ptr(), zero.data_ptr(), args.range, out.data_ptr(), nb, stream))
b.record(); torch.cuda.synchronize()
kms = a.elapsed_time(b) / args.frames
if rank == 0:
    cands = nb * side * side
    pair_bytes = cands * (2 * n * n + 4)
    uniq = nb * ((n + 2 * args.range) ** 2 + n * n) + cands * 4
    print(json.dumps({"workload": "full search r=%d, %dx%d blocks, 3840x2160 8-bit" % (args.range, n, n), "n_gpus": world,
                      "blocks_per_gpu": nb, "candidates_per_gpu": cands, "ms_per_frame": ms, "search_kernel_ms": kms,
                      "gcand_per_s_per_gpu": cands / kms / 1e6, "pair_equiv_gbps": pair_bytes / kms / 1e6,
                      "unique_gbps": uniq / kms / 1e6, "exchange_ms": ms - kms if world > 1 else 0.0,
                      "frames_per_s": 1e3 / ms}))
if world > 1:
    dist.destroy_process_group()
