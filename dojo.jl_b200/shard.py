"""Multi-GPU sharding of the environment batch (BASELINE.json north_star / SURVEY.md §8e).

Environments are independent, so the batch is split contiguously over the ranks (one process per GPU, launched with
torchrun); nothing is exchanged inside the solve.  The only collective is one all-gather of the next-state buffers per
step, for callers that need the whole batch on every rank (e.g. a centralised policy).  The helpers below are backend
agnostic (NCCL on GPUs, gloo in the CPU tests).
"""
from typing import Tuple

import numpy as np


def shard_bounds(batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of range(batch) over `world` ranks: the first (batch % world) ranks get one extra environment."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_states(z_local, batch: int, group=None):
    """All-gather the per-rank next-state shards [B_local, nz] into [batch, nz] (torch tensors, any backend).
    Handles uneven shards by padding to the largest shard."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [shard_bounds(batch, world, r) for r in range(world)]
    bmax = max(hi - lo for lo, hi in sizes)
    nz = z_local.shape[1]
    pad = torch.zeros((bmax, nz), dtype=z_local.dtype, device=z_local.device)
    pad[: z_local.shape[0]] = z_local
    out = torch.empty((world * bmax, nz), dtype=z_local.dtype, device=z_local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * bmax: r * bmax + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], dim=0)


def scatter_check(batch: int, world: int) -> np.ndarray:
    """Owner rank of every environment (for tests / diagnostics)."""
    owner = np.empty(batch, dtype=np.int32)
    for r in range(world):
        lo, hi = shard_bounds(batch, world, r)
        owner[lo:hi] = r
    return owner
