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


class StateGather:
    """The one exchange of the path (every rank receives the next states of the whole batch), for one process per GPU.

    fused = True : dojo_step_gather_async -- the step kernel writes every environment's next state straight into the gathered buffer
                   of every rank (peer memory mapped with CUDA IPC over NVLink / NVSwitch), overlapping the exchange with the solve
                   and its tail; a small wait kernel closes the step.  The IPC descriptors are exchanged once with torch.distributed.
    fused = False: fallback when peer mapping is unavailable (or DOJO_B200_GATHER=nccl): one NCCL all-gather per step on a side
                   stream behind the kernel.
    """

    def __init__(self, stepper, B_local, nz, rank, world, dist, device):
        import os
        import torch
        self.stepper, self.B, self.nz, self.rank, self.world, self.dist, self.dev = stepper, B_local, nz, rank, world, dist, device
        self.g = None
        self.fused = False
        self.why = ""
        ok = torch.zeros(1, dtype=torch.int32, device=device)
        if os.environ.get("DOJO_B200_GATHER", "p2p") != "nccl":
            try:
                self.g = stepper.gather_create(world, rank, B_local)
                mine = torch.frombuffer(bytearray(stepper.gather_export(self.g)), dtype=torch.uint8).to(device)
                allh = torch.empty(world * 128, dtype=torch.uint8, device=device)
                dist.all_gather_into_tensor(allh, mine)
                stepper.gather_connect(self.g, bytes(allh.cpu().numpy().tobytes()))
                ok += 1
            except Exception as ex:  # e.g. CUDA IPC not permitted in this container
                self.why = f"{type(ex).__name__}: {ex}"
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # all ranks or none
        self.fused = bool(ok.item() == 1)
        if self.fused:
            # the library alternates between two halves of the gathered buffer (dojo_gather_buffer = the half of the most recent step)
            self._views = {}
            self._refresh()
        else:
            self.Zall = torch.empty((world * B_local, nz), dtype=torch.float64, device=device)
            self.side = torch.cuda.Stream(device=device)
            self.ev = torch.cuda.Event()
            self.done = torch.cuda.Event()

    def _refresh(self):
        ptr = int(self.stepper.gather_buffer(self.g))
        if ptr not in self._views:
            self._views[ptr] = _wrap_device_buffer(ptr, (self.world * self.B, self.nz), self.dev)
        self.Zall = self._views[ptr]  # gathered states of the most recent step (valid until the step after the next one is issued)

    def step(self, dZ, dU, dZn, opts, dstatus=None, diters=None, stream=0):
        self.stepper.step_gather_device(self.g, dZ, dU, dZn, self.B, opts, dstatus=dstatus, diters=diters, stream=stream)
        self._refresh()

    def step_grad(self, dZ, dU, dZn, dFz, dFu, opts, dstatus=None, diters=None, stream=0):
        self.stepper.step_grad_gather_device(self.g, dZ, dU, dZn, dFz, dFu, self.B, opts, dstatus=dstatus, diters=diters, stream=stream)
        self._refresh()

    def exchange(self, z_local, stream):
        """NCCL fallback: all-gather behind the kernel; the launching stream waits for it (the step ends with the gathered buffer filled)."""
        import torch
        self.ev.record(stream)
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.ev)
            self.dist.all_gather_into_tensor(self.Zall, z_local)
            self.done.record(self.side)
        stream.wait_event(self.done)

    def exchange_host(self, z_local_host, z_all_host):
        """End-to-end leg of bench.py: the local next states are on the host already; gather them on the device and bring the whole
        batch to the host of every rank."""
        import torch
        zl = torch.from_numpy(z_local_host).to(self.dev, non_blocking=True)
        tmp = self.Zall if not self.fused else torch.empty_like(self.Zall)
        self.dist.all_gather_into_tensor(tmp, zl)
        torch.from_numpy(z_all_host).copy_(tmp, non_blocking=True)
        torch.cuda.synchronize()

    def describe(self):
        return {"kind": "fused peer writes (CUDA IPC over NVLink, dojo_step_gather_async)" if self.fused else "NCCL all_gather_into_tensor behind the kernel (side stream)",
                "bytes_per_rank_per_step": 8 * self.B * self.nz * (self.world - 1), "fallback_reason": self.why or None}

    def close(self):
        if self.g is not None:
            self.stepper.gather_destroy(self.g)
            self.g = None


def _wrap_device_buffer(ptr, shape, device):
    """torch view of a device buffer owned by the library (no copy): __cuda_array_interface__."""
    import torch

    class _Buf:
        pass
    b = _Buf()
    b.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 3, "strides": None}
    return torch.as_tensor(b, device=device)
