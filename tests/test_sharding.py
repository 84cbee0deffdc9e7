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


def test_fused_gather_kernel_side_on_the_emulation():
    """dojo_step_gather_async, kernel side (the CUDA IPC mapping and the wait kernel need GPUs: tests/test_zz_gpu_gather.py): two
    "ranks" step their shards on the CPU emulation of the step kernel; each rank's epilogue writes into BOTH gathered buffers, every
    CTA counts itself in on both counters.  Afterwards both buffers hold [shard 0 | shard 1] and both counters the number of CTAs."""
    import dojo_jl_b200 as dj
    from conftest import jittered_states, random_inputs
    from hostemu.harness import HostEmu
    mech = dj.get_mechanism("ant")
    rng = np.random.default_rng(61)
    B, world, grid = 5, 2, 2
    emu = HostEmu(mech)
    bufs = [np.full((world * B, mech.nz), np.nan) for _ in range(world)]
    flags = [np.zeros(1, dtype=np.uint64) for _ in range(world)]
    outs = []
    for r in range(world):
        Z, U = jittered_states(mech, B, rng), random_inputs(mech, B, rng)
        Zn, st, it = emu.step_gather(Z, U, r, bufs, flags, slots=2, grid=grid)
        Zref, st2, it2, _ = emu.step(Z, U, slots=2, grid=grid)
        assert np.array_equal(Zn, Zref) and np.array_equal(it, it2)  # the exchange does not change the step
        outs.append(Zn)
    full = np.concatenate(outs)
    for r in range(world):
        assert np.array_equal(bufs[r], full)
        assert int(flags[r][0]) == world * grid
