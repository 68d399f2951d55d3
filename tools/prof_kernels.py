"""Launch each hot kernel of the ctu_pass twice on bench-sized inputs (for ncu captures).
  ncu --set full --clock-control none --import-source on -k regex:'rough_cost|fused_|satd_kernel|sad_ssd' \
      -o gpurun_out/prof python tools/prof_kernels.py
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
import uvg266_b200 as U

U.init(0)
L = U.lib()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
rows = bench.rank_rows(0, 1)
if len(sys.argv) > 1:
    rows = rows[: int(sys.argv[1])]
plane_h = bench.synth_plane_rows(rows)
n_ctu = len(rows) * bench.CTU_COLS
cfg = U.QuantCfg.make(bench.QP, 8, 1, 0, 0)
reps = 2
for (n, cnt, k) in bench.ROUGH:
    nb = n_ctu * cnt
    o = torch.from_numpy(bench.gather_blocks(plane_h, n, cnt)).to(dev)
    p = o.view(nb, 1, n * n).expand(nb, k, n * n).to(torch.int16)
    p = (p + torch.randint(-24, 25, p.shape, device=dev, dtype=torch.int16)).clamp_(0, 255).to(torch.uint8).contiguous()
    out = torch.zeros(nb * k, dtype=torch.int32, device=dev)
    A, B = U.Blocks.dense(o.data_ptr(), n, n), U.Blocks.dense(p.data_ptr(), n, n)
    for _ in range(reps):
        U.check(L.uvg_b200_rough_cost_dev(8, n, C.byref(A), C.byref(B), out.data_ptr(), nb, k, stream))
for (n, cnt) in bench.TUS:
    nt = n_ctu * cnt
    r = torch.from_numpy(bench.gather_blocks(plane_h, n, cnt)).to(dev)
    p = (r.to(torch.int16) + torch.randint(-24, 25, r.shape, device=dev, dtype=torch.int16)).clamp_(0, 255).to(torch.uint8)
    rec = torch.zeros_like(r); coeff = torch.zeros(nt * n * n, dtype=torch.int16, device=dev)
    has = torch.zeros(nt, dtype=torch.uint8, device=dev)
    ssd = torch.zeros(nt, dtype=torch.int32, device=dev); cost = torch.zeros(nt, dtype=torch.int32, device=dev)
    R, P, O = (U.Blocks.dense(t.data_ptr(), n, n) for t in (r, p, rec))
    for _ in range(reps):
        U.check(L.uvg_b200_quantize_residual_dev(C.byref(cfg), 1, 0, n, n, 0, 0, 0, C.byref(R), C.byref(P), C.byref(O),
                                                 coeff.data_ptr(), has.data_ptr(), ssd.data_ptr(), cost.data_ptr(),
                                                 bench.COST_WEIGHTS, nt, stream))
n, N = 64, 1 << 15
a = torch.randint(0, 256, (N * n * n,), dtype=torch.uint8, device=dev)
b = torch.randint(0, 256, (N * n * n,), dtype=torch.uint8, device=dev)
out = torch.zeros(N, dtype=torch.int32, device=dev)
A, B = U.Blocks.dense(a.data_ptr(), n, n), U.Blocks.dense(b.data_ptr(), n, n)
for metric in (U.SAD_NXN, U.SATD):
    for _ in range(reps):
        U.check(L.uvg_b200_distortion_dev(metric, 8, n, n, C.byref(A), C.byref(B), out.data_ptr(), N, 1, stream))
torch.cuda.synchronize()
print("ok")
