#!/bin/bash
mkdir -p gpurun_out
L=dojo.jl_b200/libdojo_b200.so
{
echo "== timing (new library: LDS plan, early exit, kept limit dual in the joint node)"
DJ_ROLLOUT=0 python tools/time_variant.py $L ant 4096 10
DOJO_B200_GENERIC_PLAN=1 DJ_ROLLOUT=0 python tools/time_variant.py $L ant 4096 10
python tools/time_variant.py $L quadruped 4096 10
python tools/time_variant.py $L atlas 4096 5
echo "== parity on benchmarked states"
python tools/parity_bench_states.py ant 4096 1024 23
python tools/parity_bench_states.py quadruped 4096 512 10
echo "== gpu tests"
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
} > gpurun_out/r2_exp2.log 2>&1
tail -50 gpurun_out/r2_exp2.log
