#!/usr/bin/env python
"""Device-timed throughput of the paths added around the step, ant batch 4096 by default:
map Jacobians, get_minimal_gradients!, the environment step.  CUDA events on the launching stream, inputs resident in HBM.
Writes gpurun_out/widened_<mech>.json and prints it.   python tools/bench_widened.py [--mech ant] [--batch 4096] [--steps 5]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dojo_jl_b200 as dj  # noqa: E402
from dojo_jl_b200 import capi, environments as E  # noqa: E402
from dojo_jl_b200.solver import BatchedStepper  # noqa: E402
from bench import synthetic_batch, random_inputs, SCALE  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mech", default="ant")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    import torch
    mech = dj.get_mechanism(args.mech)
    B, K, W = args.batch, args.steps, args.warmup
    st = BatchedStepper(mech, B)
    opts = capi.solver_options()
    Z, rng = synthetic_batch(mech, B, 1234)
    U = random_inputs(mech, rng, 20 + K + W, B, SCALE[args.mech])
    for t in range(20):  # roll-in: contacts develop
        Z, _, _ = st.step(Z, U[t])
    stream = torch.cuda.current_stream().cuda_stream
    nm, ns, nu = 2 * mech.nu, 12 * mech.Nb, mech.nu
    dZ = torch.from_numpy(Z).cuda()
    dX = torch.empty((B, nm), dtype=torch.float64, device="cuda")
    st.maximal_to_minimal_device(dZ.data_ptr(), dX.data_ptr(), B, stream=stream)
    dU = torch.from_numpy(U[20:]).cuda()
    out = {"mech": args.mech, "batch": B, "steps": K, "warmup": W}
    torch.cuda.synchronize()  # dojo_minimal_gradients runs on the handle's own stream

    def timed(fn):
        for i in range(W):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            fn(W + i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / K

    dM = torch.empty((B, ns, nm), dtype=torch.float64, device="cuda")
    ms = timed(lambda i: st.maximal_to_minimal_jacobian_device(dZ.data_ptr(), dM.data_ptr(), B, stream=stream))
    out["maximal_to_minimal_jacobian"] = {"ms": ms, "env_per_s": B / ms * 1e3, "GBps_out": B * ns * nm * 8 / ms / 1e6}
    ms = timed(lambda i: st.minimal_to_maximal_jacobian_device(dZ.data_ptr(), dM.data_ptr(), B, stream=stream))
    out["minimal_to_maximal_jacobian"] = {"ms": ms, "env_per_s": B / ms * 1e3, "GBps_out": B * ns * nm * 8 / ms / 1e6}
    del dM

    dXn = torch.empty_like(dX)
    dGx = torch.empty((B, nm, nm), dtype=torch.float64, device="cuda")
    dGu = torch.empty((B, nu, nm), dtype=torch.float64, device="cuda")
    dst = torch.empty(B, dtype=torch.int32, device="cuda")
    ms = timed(lambda i: st.minimal_gradients_device(dX.data_ptr(), dU[i].data_ptr(), dXn.data_ptr(), dGx.data_ptr(), dGu.data_ptr(), B, opts,
                                                     dstatus=dst.data_ptr()))
    out["get_minimal_gradients"] = {"ms": ms, "env_steps_per_s": B / ms * 1e3, "failed": int((dst != 0).sum().item())}

    # environment step (AntARS / QuadrupedSampling / Pendulum spec of the mechanism), state carried from step to step
    cls = {"ant": E.AntARS, "quadruped": E.QuadrupedSampling, "pendulum": E.Pendulum}.get(args.mech)
    if cls is not None:
        spec = capi.env_spec(**cls.spec_kwargs)
        nse, na = st.env_sizes(spec)
        S = [torch.zeros((B, nse), dtype=torch.float64, device="cuda") for _ in range(2)]
        S[0][:, :nm] = dX
        dA = dU[:, :, spec.n_unactuated:].contiguous()
        dR = torch.empty(B, dtype=torch.float64, device="cuda")
        dD = torch.empty(B, dtype=torch.int32, device="cuda")

        def env(i):
            st.env_step_device(spec, S[i % 2].data_ptr(), dA[i].data_ptr(), S[(i + 1) % 2].data_ptr(), B, opts, dreward=dR.data_ptr(),
                               ddone=dD.data_ptr(), dstatus=dst.data_ptr(), stream=stream)
        ms = timed(env)
        out["env_step"] = {"ms": ms, "env_steps_per_s": B / ms * 1e3, "done": int(dD.sum().item()), "failed": int((dst != 0).sum().item()),
                           "mean_reward": float(dR.mean().item())}
    out["launches"] = st.launch_count
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"widened_{args.mech}.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
