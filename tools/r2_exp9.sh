#!/bin/bash
# round 2, call: two lanes per joint in set_entries! (eval_joint_pair) on / off; atlas with 4 warps per environment; GPU suite
mkdir -p gpurun_out
{
echo "== pair on / off"
timeout 300 python tools/prof_one.py ant 4096 8 fwd
DOJO_B200_NO_JOINT_PAIR=1 timeout 300 python tools/prof_one.py ant 4096 8 fwd
timeout 300 python tools/prof_one.py ant 4096 5 grad
timeout 300 python tools/prof_one.py quadruped 8192 5 fwd
DOJO_B200_NO_JOINT_PAIR=1 timeout 300 python tools/prof_one.py quadruped 8192 5 fwd
timeout 300 python tools/prof_one.py quadruped 8192 5 grad
echo "== bitwise: checksum of the state after the timed steps must not depend on the pairing"
DJ_ROLLOUT=0 timeout 300 python tools/time_variant.py dojo.jl_b200/libdojo_b200.so ant 4096 10 | cut -c1-200
DOJO_B200_NO_JOINT_PAIR=1 DJ_ROLLOUT=0 timeout 300 python tools/time_variant.py dojo.jl_b200/libdojo_b200.so ant 4096 10 | cut -c1-200
echo "== atlas: 2 vs 4 warps per environment"
timeout 300 python tools/prof_one.py atlas 4096 3 fwd
DOJO_B200_WARPS=4 timeout 300 python tools/prof_one.py atlas 4096 3 fwd
echo "== gpu tests (whole suite)"
rm -f gpurun_out/parity_stats.jsonl
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -8
} > gpurun_out/r2_exp9.log 2>&1
grep -v "config:" gpurun_out/r2_exp9.log | tail -c 5000
