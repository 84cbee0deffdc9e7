#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched Dojo step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--mech ant|quadruped|atlas|pendulum] [--batch B] [--mode fwd|grad]

Default workload = BASELINE.json configs[1]: Ant (DojoEnvironments defaults), batch 4096, forward-only step!, 1 GPU.
A "step" is one step! of every environment of the batch (the state is advanced from step to step, fresh random inputs
every step; the batch is first rolled out for 20 untimed steps so that a realistic share of environments is in contact).

One JSON line is printed by rank 0.  `value` = env-steps/s with inputs resident in HBM (CUDA events around the kernel
launches on the launching stream, max over ranks); `e2e` = the same metric through the public API with HOST buffers
(pinned staging, H2D and D2H copies inside the timed region).  `roofline` compares the algorithmic HBM bytes per step
 with the measured HBM peak; `cpu_baseline` times the CPU oracle (a port of the reference algorithm; the
reference itself is Julia and cannot run here) on the host cores.  `--impl reference` times that CPU port alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import dojo_jl_b200 as dj  # noqa: E402
from dojo_jl_b200 import capi  # noqa: E402

ALGO_BYTES = {  # per env-step, SURVEY.md §8(d): read z + read u + write z_next + status/iters (+ gradients)
    "fwd": lambda m: 8 * (2 * m.nz + m.nu) + 8,
    "grad": lambda m: 8 * (2 * m.nz + m.nu) + 8 + 8 * (12 * m.Nb) ** 2 + 8 * 12 * m.Nb * m.nu,
}
SCALE = {"ant": 1.0, "quadruped": 2.0, "atlas": 5.0, "pendulum": 1.0}


def synthetic_batch(mech, B, seed):
    """Seeded synthetic states: forward kinematics at jittered joint angles inside the limits, random base
    height / tilt, zero velocities; velocities and contacts develop during the untimed roll-in."""
    rng = np.random.default_rng(seed)
    base = mech.minimal_coordinates(mech.z0)
    n_proto = min(B, 256)
    protos = np.zeros((n_proto, mech.nz))
    for e in range(n_proto):
        coords = {}
        for j in mech.joints:
            c = np.array(base[j.name], dtype=float)
            if j.nimpulses == 0:
                c[2] += rng.uniform(-0.1, 0.25)
                c[3:6] += rng.normal(0.0, 0.15, 3)
            elif j.input_dimension > 0:
                c = c + rng.uniform(-0.3, 0.3, c.shape)
                if j.rot.nlimits:
                    lo, hi = j.rot.limit_lo, j.rot.limit_hi
                    c[j.tra.nfree:] = np.clip(c[j.tra.nfree:], lo + 0.05 * (hi - lo), hi - 0.05 * (hi - lo))
            coords[j.name] = c
        protos[e] = mech.forward_kinematics(coords)
    Z = protos[rng.integers(0, n_proto, B)].copy()
    return Z, rng


def random_inputs(mech, rng, T, B, scale):
    U = rng.uniform(-scale, scale, (T, B, mech.nu))
    off = 0
    for j in mech.joints:
        if j.nimpulses == 0:
            U[:, :, off:off + j.input_dimension] = 0.0
        off += j.input_dimension
    return U


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_throughput(mech, Z, U, opts, threads):
    """env-steps/s of the CPU oracle (oracle/, a port of the reference algorithm) on `threads` host threads (C++
    std::thread pool inside the oracle library: one mechanism instance per thread, contiguous slices of the batch)."""
    from oracle import oracle as orc  # test infrastructure: only used as the timed CPU baseline
    t0 = time.perf_counter()
    out = orc.step_batch_threads(mech, Z, U, opts, max(1, min(threads, Z.shape[0])))
    dt = time.perf_counter() - t0
    cpu_port_throughput.last = out
    return Z.shape[0] / dt, dt


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(mech_name, mode):
    p = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(f"{mech_name}_{mode}", {}).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mech", default="ant")
    ap.add_argument("--batch", type=int, default=4096, help="environments per GPU")
    ap.add_argument("--mode", default="fwd", choices=["fwd", "grad"])
    ap.add_argument("--rollin", type=int, default=20)
    ap.add_argument("--cpu-threads", type=int, default=0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    mech = dj.get_mechanism(args.mech)
    B = args.batch
    opts = capi.solver_options()
    metric = "env-steps/sec (forward step!)" if args.mode == "fwd" else "env-steps/sec (forward + IFT gradients)"
    config = {"workload": f"{args.mech} (DojoEnvironments defaults, h={mech.timestep}) batch={B}/GPU {'forward-only step!' if args.mode == 'fwd' else 'step! + get_maximal_gradients'}",
              "mechanism": args.mech, "batch_per_gpu": B, "global_batch": B * world, "mode": args.mode,
              "solver": "rtol=1e-6 btol=1e-4 max_iter=50 max_ls=10 (reference defaults)",
              "l2": "L2 flushed (256 MiB write) before every timed step", "sharding": f"env batch split over {world} GPU(s)"}
    threads = args.cpu_threads or (os.cpu_count() or 1)

    # ------------------------------------------------------------------ reference arm: CPU port on the host cores
    if args.impl == "reference":
        if rank != 0:
            return
        Z, rng = synthetic_batch(mech, B, 0xD0D0 + 1)
        U = random_inputs(mech, rng, args.warmup + args.steps, B, SCALE.get(args.mech, 1.0))
        bound = min(B, 4096)
        for t in range(min(args.rollin, 20)):  # roll-in on a bounded slice, replicated
            Z[:bound], _, _ = cpu_port_rollin(mech, Z[:bound], U[0][:bound], opts, threads)
        Z = Z[np.arange(B) % bound]
        times = []
        for t in range(args.warmup + args.steps):
            v, dt = cpu_port_throughput(mech, Z[:bound], U[t][:bound], opts, threads)
            if t >= args.warmup:
                times.append(dt)
        ms = 1e3 * float(np.mean(times)) * (B / bound)
        value = B / (ms * 1e-3)
        line = {"impl": "reference", "metric": metric, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": threads, "kind": "port",
                                 "sample": f"{bound} environments x 1 step! per timed step on {threads} host threads (C++ port of the Julia reference; Julia is not installed)"},
                "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from dojo_jl_b200.solver import BatchedStepper
    stepper = BatchedStepper(mech, B, device=local_rank)
    Z0, rng = synthetic_batch(mech, B, 0xD0D0 + 1 + 1000 * rank)
    T = args.warmup + args.steps
    U_host = random_inputs(mech, rng, args.rollin + T, B, SCALE.get(args.mech, 1.0))
    dev = torch.device("cuda", local_rank)
    Za = torch.from_numpy(Z0).to(dev)
    Zb = torch.empty_like(Za)
    U = torch.from_numpy(U_host).to(dev)
    status = torch.zeros(B, dtype=torch.int32, device=dev)
    iters = torch.zeros(B, dtype=torch.int32, device=dev)
    ng = 12 * mech.Nb
    if args.mode == "grad":
        Fz = torch.empty((B, ng, ng), dtype=torch.float64, device=dev)
        Fu = torch.empty((B, mech.nu, ng), dtype=torch.float64, device=dev)
    Zall = torch.empty((world * B, mech.nz), dtype=torch.float64, device=dev) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def launch(t, zin, zout):
        if args.mode == "fwd":
            stepper.step_device(zin.data_ptr(), U[t].data_ptr(), zout.data_ptr(), B, opts, dstatus=status.data_ptr(), diters=iters.data_ptr(),
                                stream=stream.cuda_stream)
        else:
            stepper.step_grad_device(zin.data_ptr(), U[t].data_ptr(), zout.data_ptr(), Fz.data_ptr(), Fu.data_ptr(), B, opts,
                                     dstatus=status.data_ptr(), diters=iters.data_ptr(), stream=stream.cuda_stream)
        if world > 1:  # one all-gather of the next-state buffers per step (BASELINE.json north_star)
            dist.all_gather_into_tensor(Zall, zout)

    # roll-in (untimed, forward only) so that contacts are active
    for t in range(args.rollin):
        stepper.step_device(Za.data_ptr(), U[t].data_ptr(), Zb.data_ptr(), B, opts, stream=stream.cuda_stream)
        Za, Zb = Zb, Za
    torch.cuda.synchronize()
    Uoff = args.rollin
    for t in range(args.warmup):
        launch(Uoff + t, Za, Zb)
        Za, Zb = Zb, Za
    torch.cuda.synchronize()
    Z_timed_start = Za.clone()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = stepper.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    it_sum = 0.0
    fails = 0
    for k in range(args.steps):
        flush.fill_(k & 0xFF)  # L2 flush, outside the event pair
        ev[k][0].record(stream)
        launch(Uoff + args.warmup + k, Za, Zb)
        ev[k][1].record(stream)
        Za, Zb = Zb, Za
        it_sum += float(iters.float().mean().item())
        fails += int((status != 0).sum().item())
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = stepper.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    if world > 1:
        tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        total_ms = float(tmax.item())
    ms_per_step = total_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ---- fused rollout (dojo_rollout_async): the same K steps of the same batch in ONE launch, every environment advanced
    #      through all steps by the CTA slot that picked it up (no per-step launch / tail).  Reported beside the headline.
    rollout = None
    if args.mode == "fwd" and world == 1:
        Zr, Zf = Z_timed_start.clone(), torch.empty_like(Z_timed_start)
        Ur = U[Uoff + args.warmup: Uoff + T].contiguous()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stepper.rollout_device(Zr.data_ptr(), Ur.data_ptr(), Zf.data_ptr(), B, args.steps, opts, stream=stream.cuda_stream)  # warm
        torch.cuda.synchronize()
        flush.fill_(1)
        r0.record(stream)
        stepper.rollout_device(Zr.data_ptr(), Ur.data_ptr(), Zf.data_ptr(), B, args.steps, opts, stream=stream.cuda_stream)
        r1.record(stream)
        torch.cuda.synchronize()
        rms = r0.elapsed_time(r1) / args.steps
        rollout = {"value": B / (rms * 1e-3), "unit": "env-steps/s", "ms_per_step": rms, "steps_fused": args.steps, "launches": 1,
                   "final_state_matches_stepwise": bool(torch.equal(Zf, Za))}

    # ---- minimal-coordinate path: the maps either side of step! and step_minimal_coordinates! = map + step + map
    minimal = None
    if args.mode == "fwd" and world == 1 and mech.nu > 0:
        Xd = torch.empty((B, 2 * mech.nu), dtype=torch.float64, device=dev)
        Zd = torch.empty_like(Z_timed_start)
        m0, m1, m2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        stepper.maximal_to_minimal_device(Z_timed_start.data_ptr(), Xd.data_ptr(), B, stream=stream.cuda_stream)  # warm
        stepper.minimal_to_maximal_device(Xd.data_ptr(), Zd.data_ptr(), B, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        flush.fill_(2)
        m0.record(stream)
        stepper.maximal_to_minimal_device(Z_timed_start.data_ptr(), Xd.data_ptr(), B, stream=stream.cuda_stream)
        m1.record(stream)
        stepper.minimal_to_maximal_device(Xd.data_ptr(), Zd.data_ptr(), B, stream=stream.cuda_stream)
        m2.record(stream)
        torch.cuda.synchronize()
        bytes_map = 8 * B * (2 * mech.nu + mech.nz)
        t_mm, t_mM = m0.elapsed_time(m1), m1.elapsed_time(m2)
        minimal = {"maximal_to_minimal_us": 1e3 * t_mm, "minimal_to_maximal_us": 1e3 * t_mM,
                   "maximal_to_minimal_GBps": bytes_map / (t_mm * 1e-3) / 1e9, "minimal_to_maximal_GBps": bytes_map / (t_mM * 1e-3) / 1e9,
                   "round_trip_max_abs_err": float((Zd - Z_timed_start).abs().max().item()),
                   "step_minimal_ms": ms_per_step + t_mm + t_mM,
                   "step_minimal_value": B / ((ms_per_step + t_mm + t_mM) * 1e-3), "unit": "env-steps/s"}

    # ---- e2e through the public host API (pinned staging + H2D + kernel + D2H), same workload, N = 1 path per rank
    # Host buffers are page-locked (torch pin_memory), as the contract asks: the library DMAs straight from / to them.
    def pinned(shape, dtype):
        return torch.empty(shape, dtype=dtype, pin_memory=True).numpy()
    Zh, Zh2 = pinned((B, mech.nz), torch.float64), pinned((B, mech.nz), torch.float64)
    Zh[:] = Z_timed_start.cpu().numpy()
    Uh = pinned((args.steps, B, mech.nu), torch.float64)
    Uh[:] = U_host[Uoff + args.warmup: Uoff + T]
    sth, ith = pinned((B,), torch.int32), pinned((B,), torch.int32)
    if args.mode == "grad":
        Fzh, Fuh = pinned((B, ng, ng), torch.float64), pinned((B, mech.nu, ng), torch.float64)
    e2e_t = []
    # untimed warm-up of the host path (first call allocates the library's staging buffers); it does not advance the state
    if args.mode == "fwd":
        stepper.step(Zh, Uh[0], opts, out=(Zh2, sth, ith))
    else:
        stepper.step_grad(Zh, Uh[0], opts, out=(Zh2, Fzh, Fuh, sth, ith))
    if world > 1:
        dist.barrier()
    for k in range(min(args.steps, 10)):
        flush.fill_(k & 0xFF)  # same L2 flush as the device-timed steps, outside the timed region
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if args.mode == "fwd":
            stepper.step(Zh, Uh[k], opts, out=(Zh2, sth, ith))
            loss = float(Zh2[0, 2])  # the result is on the host
        else:
            stepper.step_grad(Zh, Uh[k], opts, out=(Zh2, Fzh, Fuh, sth, ith))
            loss = float(Fzh[0, 0, 0])
        e2e_t.append(time.perf_counter() - t0)
        Zh, Zh2 = Zh2, Zh
    e2e_ms = 1e3 * float(np.mean(e2e_t))
    if world > 1:
        tmax = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_ms = float(tmax.item())
    h2d = 8 * B * (mech.nz + mech.nu)
    d2h = 8 * B * mech.nz + 8 * B + (8 * B * (ng * ng + ng * mech.nu) if args.mode == "grad" else 0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant (only) kernel
    peak, peak_src = peak_hbm()
    algo = ALGO_BYTES[args.mode](mech) * B
    kernel_ms = float(np.mean(step_ms))
    achieved = algo / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(args.mech, args.mode),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": algo, "kernel_ms": kernel_ms,
                "note": "the KKT system lives in shared memory: the kernel is latency bound (dependent fp64 block algebra, 8 warps/SM), not HBM bound (DESIGN.md)"}
    # ---- CPU baseline (oracle port) on a bounded sample
    cpu = None
    try:
        bound = min(B, 2048)
        Zs = Z_timed_start[:bound].cpu().numpy()
        Us = U_host[Uoff + args.warmup][:bound]
        v_all, dt_all = cpu_port_throughput(mech, Zs, Us, opts, threads)
        v_one, dt_one = cpu_port_throughput(mech, Zs[:256], Us[:256], opts, 1)
        cpu = {"value": v_all, "unit": "env-steps/s", "cores": threads, "kind": "port",
               "sample": f"{bound} environments x 1 step! of the same batch on {threads} threads ({dt_all:.2f} s); single thread: {v_one:.0f} env-steps/s on 256 environments",
               "single_thread_value": v_one}
    except Exception as ex:  # the oracle is test infrastructure; its absence must not break the product bench
        cpu = {"value": None, "unit": "env-steps/s", "cores": 0, "kind": "port", "sample": f"unavailable: {ex}"}
    line = {"metric": metric, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "rollout": rollout, "minimal_coordinates": minimal, "mean_newton_iters": it_sum / args.steps, "failed_env_steps": fails,
            "shared_bytes_per_env": stepper.shared_bytes_per_env}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_port_rollin(mech, Z, U, opts, threads):
    from oracle import oracle as orc
    Zn, st, it = orc.step_batch_threads(mech, Z, U, opts, max(1, min(threads, Z.shape[0])))
    return Zn, st, it


if __name__ == "__main__":
    main()
