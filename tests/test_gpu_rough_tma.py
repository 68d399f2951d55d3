"""GPU parity of the tcgen05 + tensor-map-TMA rough-cost kernel (rough_tma.cu): dense originals and candidates, every
candidate's min(SATD, 2 SAD) against the oracle (satd_NxN / sad_NxN restatements pinned on the reference), incl. partial
last tiles, candidates-per-block counts that straddle tile boundaries in every phase, and a picture-sized batch against
the packed-ALU kernel. The kernel is selected by UVG266_B200_ROUGH=tc5 (process-wide), so the body runs in a child."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

CHILD = r'''
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(%(here)r, ".."))
sys.path.insert(0, %(here)r)
import uvg266_b200 as U
from _libs import load_oracle, ptr
U.init(0)
L = U.lib(); O = load_oracle(8)
rng = np.random.default_rng(77)
def run(n, k, nb, check_all):
    orig = rng.integers(0, 256, (nb, n * n), dtype=np.uint8)
    noise = rng.integers(-40, 41, (nb, k, n * n))
    cand = np.clip(orig[:, None, :].astype(int) + noise * (1 + np.arange(k))[None, :, None] // 4, 0, 255).astype(np.uint8)
    cand[0, 0] = orig[0]                      # one exact match: cost 0
    d_o = torch.from_numpy(orig).cuda(); d_c = torch.from_numpy(cand).cuda()
    out = torch.full((nb * k,), -1, dtype=torch.int32, device="cuda")
    A = U.Blocks.dense(d_o.data_ptr(), n, n); B = U.Blocks.dense(d_c.data_ptr(), n, n)
    before = L.uvg_b200_rough_tma_launches()
    U.check(L.uvg_b200_rough_cost_dev(8, n, C.byref(A), C.byref(B), out.data_ptr(), nb, k, None))
    torch.cuda.synchronize()
    used = L.uvg_b200_rough_tma_launches() - before
    got = out.cpu().numpy().astype(np.uint32).reshape(nb, k)
    blocks = range(nb) if check_all else list(range(0, nb, max(1, nb // 97))) + [nb - 1]
    for b in blocks:
        o = np.ascontiguousarray(orig[b])
        for c in range(k):
            p = np.ascontiguousarray(cand[b, c])
            e = min(O.orc_satd_nxn(n, ptr(o), ptr(p)), 2 * O.orc_sad_nxn(n, ptr(o), ptr(p)))
            assert got[b, c] == e, (n, k, nb, b, c, int(got[b, c]), e)
    return used, got, (d_o, d_c, A, B)
for (n, k, nb) in ((8, 23, 2), (8, 23, 520), (8, 24, 10), (8, 10, 64), (8, 67, 38), (8, 31, 1002),
                   (8, 15, 46), (8, 17, 26), (8, 20, 122), (8, 22, 14), (8, 21, 1000),      # block-aligned tiles: every rows-per-block count, odd / even k, partial last tiles

                   (16, 24, 1), (16, 24, 131), (16, 10, 57), (16, 67, 9), (16, 23, 300)):
    used, _, _ = run(n, k, nb, nb <= 64)
    assert used == 1, ("the tcgen05 + TMA kernel did not take", n, k, nb)
# layouts it must decline (fall through to the other kernels, still correct)
used, _, _ = run(8, 23, 7, True)            # odd number of 8x8 blocks
assert used == 0
used, _, _ = run(8, 4, 16, True)            # too few candidates per block for 16 original rows
assert used == 0
# one 4K picture's worth of 8x8 and 16x16 candidates against the packed-ALU kernel (separate process default)
for (n, k, nb) in ((8, 23, 2040 * 64), (16, 24, 2040 * 16)):
    used, got, keep = run(n, k, nb, False)
    assert used == 1
    np.save(os.path.join(%(tmp)r, "tc5_%%d.npy" %% n), got)
print("CHILD_OK")
'''


@pytest.mark.parametrize("mode", ["tc5", None, "flat"])
def test_rough_tma_kernel_vs_oracle(tmp_path, mode):
    """mode None: no UVG266_B200_ROUGH at all - the kernel must be what the library picks by default; "flat": the 8x8
    levels with 15..24 candidates stay on the flat tile layout instead of the block-aligned one."""
    env = dict(os.environ)
    env.pop("UVG266_B200_ROUGH", None)
    env.pop("UVG266_B200_ROUGH_TILES", None)
    if mode == "flat":
        env["UVG266_B200_ROUGH_TILES"] = "flat"
    elif mode:
        env["UVG266_B200_ROUGH"] = mode
    code = CHILD % {"here": HERE, "tmp": str(tmp_path)}
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "CHILD_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
