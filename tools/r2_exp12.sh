#!/bin/bash
# round 2: work-queue key with / without the previous-iterations term; forward share of the SMs in dojo_step_grad
mkdir -p gpurun_out
{
for m in 1 3; do
echo "== DOJO_B200_LPT=$m"
DOJO_B200_LPT=$m DJ_ROLLOUT=0 timeout 300 python tools/time_variant.py dojo.jl_b200/libdojo_b200.so ant 4096 10 | cut -c1-200
DOJO_B200_LPT=$m timeout 300 python tools/prof_one.py ant 4096 8 fwd
DOJO_B200_LPT=$m timeout 300 python tools/prof_one.py quadruped 8192 5 fwd
DOJO_B200_LPT=$m timeout 300 python tools/prof_one.py atlas 4096 3 fwd
done
for f in 1.0 0.85 0.75 0.65 0.55; do
echo "== DOJO_B200_GRAD_FWD_FRAC=$f"
DOJO_B200_GRAD_FWD_FRAC=$f timeout 300 python tools/prof_one.py ant 4096 5 grad
DOJO_B200_GRAD_FWD_FRAC=$f timeout 300 python tools/prof_one.py quadruped 8192 5 grad
done
DOJO_B200_GRAD_FWD_FRAC=0.75 timeout 300 python tools/prof_one.py atlas 1024 2 grad
DOJO_B200_GRAD_FWD_FRAC=1.0 timeout 300 python tools/prof_one.py atlas 1024 2 grad
DOJO_B200_GRAD_FWD_FRAC=0.75 timeout 300 python tools/prof_one.py ant 8192 5 grad
DOJO_B200_GRAD_FWD_FRAC=1.0 timeout 300 python tools/prof_one.py ant 8192 5 grad
} > gpurun_out/r2_exp12.log 2>&1
grep -v "config:" gpurun_out/r2_exp12.log | tail -c 5000
