"""Reference-picture exchange for CTU-row-band sharding across GPUs.

All-intra frames need no communication. For inter frames every GPU must see the whole
reconstructed reference picture before the next frame's motion search, while each GPU only
reconstructs its own band of CTU rows. That is exactly one collective per coded frame: an
all-gather of the bands (equivalently one broadcast per band, as the north star words it) over
NCCL / NVLink - 4K 8-bit luma+chroma is ~12.5 MB, i.e. tens of microseconds of wire time.

One process per GPU, torch.distributed for the plumbing (backend "nccl" on GPUs, "gloo" in the
CPU tests). Bands are balanced so that the partition is identical on every rank.
"""
import torch
import torch.distributed as dist

CTU = 64


def band_rows(rank, world, ctu_rows):
    """CTU rows [lo, hi) owned by `rank` (balanced partition, same rule as bench.py)."""
    return ctu_rows * rank // world, ctu_rows * (rank + 1) // world


def exchange_reference(plane, rank=None, world=None):
    """In place: every rank has written pixel rows of its own band into `plane`
    ([H, stride] uint8/int16 tensor, H a multiple of 64); on return every rank holds all bands.

    Bands can differ by one CTU row, so the gather is done on equal-size padded slots and the
    padding is dropped - still a single collective."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    if world == 1:
        return plane
    H, stride = plane.shape
    ctu_rows = H // CTU
    slot = (ctu_rows + world - 1) // world * CTU              # rows per slot
    lo, hi = band_rows(rank, world, ctu_rows)
    send = torch.zeros((slot, stride), dtype=plane.dtype, device=plane.device)
    send[: (hi - lo) * CTU] = plane[lo * CTU: hi * CTU]
    recv = torch.empty((world * slot, stride), dtype=plane.dtype, device=plane.device)
    dist.all_gather_into_tensor(recv, send)
    for r in range(world):
        rlo, rhi = band_rows(r, world, ctu_rows)
        if r != rank:
            plane[rlo * CTU: rhi * CTU] = recv[r * slot: r * slot + (rhi - rlo) * CTU]
    return plane
