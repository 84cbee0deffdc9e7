#!/bin/bash
# round 2: order of the work queue -- least likely to stall last (risk key from the state) vs previous iterations vs index order
mkdir -p gpurun_out
{
for m in 1 2 0; do
echo "== DOJO_B200_LPT=$m"
DOJO_B200_LPT=$m DJ_ROLLOUT=0 timeout 300 python tools/time_variant.py dojo.jl_b200/libdojo_b200.so ant 4096 10 | cut -c1-200
DOJO_B200_LPT=$m timeout 300 python tools/prof_one.py ant 4096 8 fwd
DOJO_B200_LPT=$m timeout 300 python tools/prof_one.py ant 4096 5 grad
DOJO_B200_LPT=$m timeout 300 python tools/prof_one.py quadruped 8192 5 fwd
DOJO_B200_LPT=$m timeout 300 python tools/prof_one.py quadruped 8192 5 grad
DOJO_B200_LPT=$m timeout 300 python tools/prof_one.py atlas 4096 3 fwd
done
DOJO_B200_LPT=1 timeout 300 python tools/prof_one.py ant 8192 5 fwd
DOJO_B200_LPT=0 timeout 300 python tools/prof_one.py ant 8192 5 fwd
DOJO_B200_LPT=1 timeout 300 python tools/prof_one.py ant 2048 8 fwd
DOJO_B200_LPT=0 timeout 300 python tools/prof_one.py ant 2048 8 fwd
echo "== gpu tests"
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -5
} > gpurun_out/r2_exp11.log 2>&1
grep -v "config:" gpurun_out/r2_exp11.log | tail -c 5000
