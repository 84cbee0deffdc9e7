"""Correctness of the fused multi-GPU exchange (dojo_step_gather_async): run under torchrun with N >= 2 ranks.

Every rank steps its own shard of an ant batch T times with the exchange fused into the step kernel (peer writes into every rank's
gathered buffer) and compares, after every step, its gathered buffer with an NCCL all-gather of the per-rank results of the plain
dojo_step_async: they must be bit-identical on every rank.  Also checks the gradient variant.  Prints one line per rank.
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import bench
import dojo_jl_b200 as dj
from dojo_jl_b200.shard import StateGather
from dojo_jl_b200.solver import BatchedStepper


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    mech = dj.get_mechanism("ant")
    B, T = 1000, 6
    Z0, rng = bench.synthetic_batch(mech, B, 100 + rank, "ant")
    U = torch.from_numpy(bench.random_inputs(mech, rng, T, B, 1.0)).to(dev)
    st = BatchedStepper(mech, B, device=local)
    g = StateGather(st, B, mech.nz, rank, world, dist, dev)
    stream = torch.cuda.current_stream()
    Za, Zb, Zp = torch.from_numpy(Z0).to(dev), torch.empty((B, mech.nz), dtype=torch.float64, device=dev), torch.empty((B, mech.nz), dtype=torch.float64, device=dev)
    ref_all = torch.empty((world * B, mech.nz), dtype=torch.float64, device=dev)
    ok = g.fused
    prev_view, prev_ref = None, None
    for t in range(T):
        st.step_device(Za.data_ptr(), U[t].data_ptr(), Zp.data_ptr(), B, stream=stream.cuda_stream)          # plain step
        dist.all_gather_into_tensor(ref_all, Zp)
        if g.fused:
            g.step(Za.data_ptr(), U[t].data_ptr(), Zb.data_ptr(), None, stream=stream.cuda_stream)           # fused step + exchange
            torch.cuda.synchronize()
            ok = ok and torch.equal(Zb, Zp) and torch.equal(g.Zall, ref_all)
            # the gathered states of the PREVIOUS step are still intact (steps alternate between two halves of the buffer, so that a fast
            # rank writing step t + 1 cannot overwrite what a slow rank still reads of step t)
            if prev_view is not None:
                ok = ok and torch.equal(prev_view, prev_ref) and prev_view.data_ptr() != g.Zall.data_ptr()
            prev_view, prev_ref = g.Zall, ref_all.clone()
        else:
            Zb.copy_(Zp)
        Za, Zb = Zb, Za
    if g.fused:  # gradient variant
        ng = 12 * mech.Nb
        Fz = torch.empty((B, ng, ng), dtype=torch.float64, device=dev)
        Fu = torch.empty((B, mech.nu, ng), dtype=torch.float64, device=dev)
        g.step_grad(Za.data_ptr(), U[0].data_ptr(), Zb.data_ptr(), Fz.data_ptr(), Fu.data_ptr(), None, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        st.step_device(Za.data_ptr(), U[0].data_ptr(), Zp.data_ptr(), B, stream=stream.cuda_stream)
        dist.all_gather_into_tensor(ref_all, Zp)
        torch.cuda.synchronize()
        ok = ok and torch.equal(g.Zall, ref_all) and bool(torch.isfinite(Fz).all())
    print(f"rank {rank}/{world}: fused={g.fused} ({g.why}) gathered == all_gather: {ok}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    g.close()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
