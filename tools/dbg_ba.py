import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import uvg266_b200 as U
from _libs import load_oracle, ptr
U.init(0)
L = U.lib(); O = load_oracle(8)
rng = np.random.default_rng(5)
n, k, nb = 8, 23, 24
orig = rng.integers(0, 256, (nb, 64), dtype=np.uint8)
cand = np.clip(orig[:, None, :].astype(int) + rng.integers(-40, 41, (nb, k, 64)), 0, 255).astype(np.uint8)
d_o = torch.from_numpy(orig).cuda(); d_c = torch.from_numpy(cand).cuda()
out = torch.full((nb * k,), -1, dtype=torch.int32, device="cuda")
A = U.Blocks.dense(d_o.data_ptr(), n, n); B = U.Blocks.dense(d_c.data_ptr(), n, n)
U.check(L.uvg_b200_rough_cost_dev(8, n, C.byref(A), C.byref(B), out.data_ptr(), nb, k, None))
torch.cuda.synchronize()
got = out.cpu().numpy().astype(np.int64).reshape(nb, k)
def cost(o, p):
    o = np.ascontiguousarray(o); p = np.ascontiguousarray(p)
    return min(O.orc_satd_nxn(n, ptr(o), ptr(p)), 2 * O.orc_sad_nxn(n, ptr(o), ptr(p)))
exp = np.array([[cost(orig[b], cand[b, c]) for c in range(k)] for b in range(nb)])
print("launches", L.uvg_b200_rough_tma_launches())
for b in range(nb):
    print(b, "".join("." if got[b, c] == exp[b, c] else "x" for c in range(k)))
# for wrong ones: is it the cost of (orig b', cand (b, c))?  or sad/satd separately wrong?
bad = [(b, c) for b in range(nb) for c in range(k) if got[b, c] != exp[b, c]][:6]
for (b, c) in bad:
    alt = [bb for bb in range(nb) if cost(orig[bb], cand[b, c]) == got[b, c]]
    alt2 = [(bb, cc) for bb in range(nb) for cc in range(k) if exp[bb, cc] == got[b, c]][:4]
    print((b, c), "got", got[b, c], "exp", exp[b, c], "matches orig of block", alt, "equals exp of", alt2)
