"""Host-side logic of the N>1 path, on CPU with the gloo backend (world_size 2 and 4):
the CTU-row band sharding of bench.py covers every (frame, CTU row) exactly once, gives every
rank the same amount of work (weak scaling), and needs no data-path collective for all-intra;
the only communication is the max-over-ranks reduction of the step time."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rows = bench.rank_rows(rank, world)
    mine = torch.tensor([f * 1000 + r for (f, r) in rows], dtype=torch.int64)
    n = torch.tensor([mine.numel()], dtype=torch.int64)
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, n)
    pad = torch.full((bench.CTU_ROWS + 8,), -1, dtype=torch.int64)
    pad[: mine.numel()] = mine
    allrows = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(allrows, pad)
    # max-over-ranks of a per-rank "time", as bench.py does for the step time
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # the synthetic plane of a rank only depends on (frame, row): two ranks asked for the same
    # row must generate identical pixels
    plane = bench.synth_plane_rows(rows[:1])
    chk = torch.tensor([int(plane.astype("int64").sum())], dtype=torch.int64)
    if rank == 0:
        q.put((int(t.item()), [int(s.item()) for s in sizes],
               [[int(v) for v in a.tolist() if v >= 0] for a in allrows], int(chk.item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_band_sharding(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + world * 7 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    tmax, sizes, allrows, _ = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == world
    assert sizes == [bench.CTU_ROWS] * world                      # equal work per rank
    flat = [v for rows in allrows for v in rows]
    assert len(flat) == len(set(flat)) == world * bench.CTU_ROWS    # every (frame,row) exactly once
    assert set(flat) == {f * 1000 + r for f in range(world) for r in range(bench.CTU_ROWS)}


def test_single_rank_is_the_whole_frame():
    rows = bench.rank_rows(0, 1)
    assert rows == [(0, r) for r in range(bench.CTU_ROWS)]
    assert len(rows) * bench.CTU_COLS == 2040


def test_workload_bytes_match_design():
    rough, tus = bench.bytes_per_ctu()
    assert rough == {8: 104192, 16: 103936}
    assert tus == {4: 21360, 8: 39480, 16: 30936, 32: 25645}
