#!/bin/bash
# round 2: solve() with converged half-warps (full-warp syncs); smoke(); GPU suite
mkdir -p gpurun_out
{
echo "== smoke"
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== timing"
DJ_ROLLOUT=0 timeout 300 python tools/time_variant.py dojo.jl_b200/libdojo_b200.so ant 4096 10 | cut -c1-200
timeout 300 python tools/prof_one.py ant 4096 8 fwd
timeout 300 python tools/prof_one.py quadruped 8192 5 fwd
timeout 300 python tools/prof_one.py atlas 4096 3 fwd
timeout 300 python tools/grad_steps.py ant 4096
echo "== gpu tests"
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -5
} > gpurun_out/r2_exp14.log 2>&1
grep -v "config:" gpurun_out/r2_exp14.log | tail -c 4000
