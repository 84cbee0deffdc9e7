#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python tools/prof_one.py atlas 4096 3 fwd
DOJO_B200_WARPS=4 timeout 300 python tools/prof_one.py atlas 4096 3 fwd
DOJO_B200_WARPS=4 DOJO_B200_NO_JOINT_PAIR=1 timeout 300 python tools/prof_one.py atlas 4096 3 fwd
timeout 300 python tools/grad_steps.py atlas 1024
DOJO_B200_WARPS=4 timeout 300 python tools/grad_steps.py atlas 1024
} > gpurun_out/r2_exp15.log 2>&1
grep -v "config:" gpurun_out/r2_exp15.log
