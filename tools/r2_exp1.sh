#!/bin/bash
# round-2 experiment 1: baseline numbers on this pool + latency structure of the forward kernel
mkdir -p gpurun_out
L=dojo.jl_b200/libdojo_b200.so
P=build_variants/libdojo_prof.so
{
echo "== bench default"; python bench.py --steps 10 --warmup 3
echo "== env times (profile build)"; python tools/env_times.py $P ant 4096
echo "== lone-env latency: B=148, 1 slot per CTA, nw=2/4/8"
for w in 2 4 8; do DOJO_B200_WARPS=$w DOJO_B200_SLOTS=1 DJ_ROLLOUT=0 python tools/time_variant.py $L ant 148 10; done
echo "== co-resident latency: B=592 (4 slots), B=296 (2 slots), nw=2"
DJ_ROLLOUT=0 python tools/time_variant.py $L ant 592 10
DOJO_B200_SLOTS=2 DJ_ROLLOUT=0 python tools/time_variant.py $L ant 296 10
DOJO_B200_SLOTS=3 DJ_ROLLOUT=0 python tools/time_variant.py $L ant 444 10
echo "== full batch variants"
DOJO_B200_WARPS=4 DJ_ROLLOUT=0 python tools/time_variant.py $L ant 4096 10
DOJO_B200_SLOTS=3 DJ_ROLLOUT=0 python tools/time_variant.py $L ant 4096 10
DJ_PROF=1 DJ_ROLLOUT=0 python tools/time_variant.py $P ant 4096 10
echo "== other mechanisms"
python tools/time_variant.py $L quadruped 4096 10
python tools/time_variant.py $L atlas 4096 5
} > gpurun_out/r2_exp1.log 2>&1
tail -60 gpurun_out/r2_exp1.log
