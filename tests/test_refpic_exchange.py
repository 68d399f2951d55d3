"""Reference-picture band exchange (uvg266_b200/refpic.py) on CPU with gloo, world sizes 2 and 3
(uneven bands): after the single all-gather every rank holds the full picture."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from uvg266_b200 import refpic
    H, stride = 5 * 64, 200          # 5 CTU rows: uneven split for world 2 and 3
    g = torch.Generator().manual_seed(7)
    full = torch.randint(0, 256, (H, stride), dtype=torch.uint8, generator=g)
    plane = torch.zeros_like(full)
    lo, hi = refpic.band_rows(rank, world, H // 64)
    plane[lo * 64: hi * 64] = full[lo * 64: hi * 64]       # this rank "reconstructed" its band
    refpic.exchange_reference(plane)
    ok = torch.equal(plane, full)
    oks = [None] * world
    dist.all_gather_object(oks, bool(ok))
    if rank == 0:
        q.put(oks)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + world * 11 + (os.getpid() % 300)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    oks = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert oks == [True] * world


def test_bands_partition():
    from uvg266_b200 import refpic
    for world in (1, 2, 3, 4, 8):
        rows = []
        for r in range(world):
            lo, hi = refpic.band_rows(r, world, 34)
            rows += list(range(lo, hi))
        assert rows == list(range(34))


def test_c_abi_band_partition_matches_python():
    """uvg_b200_band_rows (the partition uvg_b200_refpic_exchange broadcasts by) is the rule of refpic.band_rows and
    of bench.py's bands; needs no GPU."""
    import ctypes as C
    import uvg266_b200 as U
    from uvg266_b200 import refpic
    L = U.lib()
    for world in (1, 2, 3, 4, 8):
        for ctu_rows in (1, 5, 34, 68):
            for r in range(world):
                lo, hi = C.c_int(), C.c_int()
                L.uvg_b200_band_rows(r, world, ctu_rows, C.byref(lo), C.byref(hi))
                assert (lo.value, hi.value) == refpic.band_rows(r, world, ctu_rows)
