"""A/B of the rough-cost kernels on the bench workload (one 4K frame: 8x8 k=23 and 16x16 k=24, dense operands).
  UVG266_B200_ROUGH=tc5 python tools/rough_ab.py"""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
import uvg266_b200 as U
U.init(0); L = U.lib()
res = {"mode": os.environ.get("UVG266_B200_ROUGH", "default")}
for (n, cnt, k) in bench.ROUGH:
    nb = 2040 * cnt
    o = torch.randint(0, 256, (nb, n * n), dtype=torch.uint8, device="cuda")
    p = (o.view(nb, 1, n * n).to(torch.int16) + torch.randint(-24, 25, (nb, k, n * n), device="cuda", dtype=torch.int16)).clamp_(0, 255).to(torch.uint8).contiguous()
    out = torch.zeros(nb * k, dtype=torch.int32, device="cuda")
    A = U.Blocks.dense(o.data_ptr(), n, n); B = U.Blocks.dense(p.data_ptr(), n, n)
    st = torch.cuda.current_stream().cuda_stream
    f = lambda: U.check(L.uvg_b200_rough_cost_dev(8, n, C.byref(A), C.byref(B), out.data_ptr(), nb, k, st))
    ms = bench._time(torch, f, iters=20, warm=3)
    nbytes = nb * ((1 + k) * n * n + 4 * k)
    res["n%d" % n] = {"ms": ms, "gbps": nbytes / ms / 1e6, "tc5_launches": int(L.uvg_b200_rough_tma_launches()), "checksum": int(out.long().sum().item())}
print(json.dumps(res))
