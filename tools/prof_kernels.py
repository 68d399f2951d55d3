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

# ---- widened rows and standalone steps (one launch each, bench-sized) --------------------------------
import numpy as np
W, H = bench.FRAME_W, bench.CTU_ROWS * 64
cur_h = bench.synth_plane_rows([(1, r) for r in range(bench.CTU_ROWS)])
d_plane, d_cur = torch.from_numpy(plane_h).to(dev), torch.from_numpy(cur_h).to(dev)
for n, modes in ((8, list(range(0, 67, 3))), (16, list(range(0, 67, 3)) + [1])):
    gp = 2 * n + 4
    ys, xs = np.meshgrid(np.arange(0, H, n), np.arange(0, W, n), indexing="ij")
    ys, xs = ys.ravel(), xs.ravel()
    i = np.arange(gp)
    top = plane_h[np.clip(ys - 1, 0, H - 1)[:, None], np.clip(xs[:, None] - 1 + i[None, :], 0, W - 1)]
    left = plane_h[np.clip(ys[:, None] - 1 + i[None, :], 0, H - 1), np.clip(xs - 1, 0, W - 1)[:, None]]
    d_refs = torch.from_numpy(np.ascontiguousarray(np.stack([top, left], axis=1))).to(dev)
    m8 = np.array(modes, dtype=np.int8)
    cost = torch.zeros(ys.size * len(modes), dtype=torch.int32, device=dev)
    desc = U.Blocks.plane(d_plane.data_ptr(), W, W, n, n)
    U.check(L.uvg_b200_intra_rough_dev(8, n, C.byref(desc), d_refs.data_ptr(), m8.ctypes.data, len(modes), cost.data_ptr(), ys.size, stream))
offs = [(0, 0)] + [(dx, dy) for dy in (-8, 0, 8) for dx in (-8, 0, 8) if (dx, dy) != (0, 0)] + \
       [(dx, dy) for dy in (-4, 0, 4) for dx in (-4, 0, 4) if (dx, dy) != (0, 0)]
offsets = np.array(offs, dtype=np.int8)
ys, xs = np.meshgrid(np.arange(0, H, 16), np.arange(0, W, 16), indexing="ij")
rng = np.random.default_rng(7)
arrs = [xs.ravel().astype(np.int32), ys.ravel().astype(np.int32), (rng.integers(-8, 9, ys.size) * 16).astype(np.int32),
        (rng.integers(-8, 9, ys.size) * 16).astype(np.int32)]
dj = [torch.from_numpy(a_).to(dev) for a_ in arrs]
cost = torch.zeros(ys.size * len(offs), dtype=torch.int32, device=dev)
U.check(L.uvg_b200_frac_search_dev(8, 16, 16, d_cur.data_ptr(), d_plane.data_ptr(), W, W, H, dj[0].data_ptr(), dj[1].data_ptr(),
                                   dj[2].data_ptr(), dj[3].data_ptr(), offsets.ctypes.data, len(offs), cost.data_ptr(), ys.size, stream))
pred = torch.empty(ys.size * 256, dtype=torch.uint8, device=dev)
U.check(L.uvg_b200_ipol_dev(8, 0, 0, d_plane.data_ptr(), W, W, H, 16, 16, dj[0].data_ptr(), dj[1].data_ptr(), dj[2].data_ptr(),
                            dj[3].data_ptr(), pred.data_ptr(), ys.size, stream))
NT = 1 << 18
src = torch.randint(-255, 256, (NT * 1024,), dtype=torch.int16, device=dev)
dst = torch.zeros_like(src)
U.check(L.uvg_b200_transform_dev(0, 8, 32, 32, 0, 0, 0, 0, src.data_ptr(), dst.data_ptr(), NT, stream))
U.check(L.uvg_b200_transform_dev(1, 8, 32, 32, 0, 0, 0, 0, dst.data_ptr(), src.data_ptr(), NT, stream))
U.check(L.uvg_b200_quant_dev(C.byref(cfg), src.data_ptr(), dst.data_ptr(), 32, 32, 0, 0, 0, NT, stream))
torch.cuda.synchronize()
print("ok (widened rows)")
