"""N > 1 path on CPU: the batch split and the per-step all-gather of next states with world_size-2 gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dojo_jl_b200 import shard


def test_shard_bounds_cover_the_batch():
    for batch in (1, 7, 64, 4096, 65537):
        for world in (1, 2, 3, 8):
            owner = shard.scatter_check(batch, world)
            assert (np.diff(owner) >= 0).all() and owner.min() == 0
            counts = np.bincount(owner, minlength=world)
            assert counts.sum() == batch and counts.max() - counts.min() <= 1


def _worker(rank, world, port, batch, nz, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.shard_bounds(batch, world, rank)
    full = torch.arange(batch * nz, dtype=torch.float64).reshape(batch, nz)
    local_next = full[lo:hi] * 2.0 + 1.0  # stand-in for the local step: any per-environment map
    gathered = shard.all_gather_states(local_next, batch)
    ok = torch.equal(gathered, full * 2.0 + 1.0)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [10, 7])
def test_all_gather_states_gloo_world2(batch):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, batch, 13, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]
