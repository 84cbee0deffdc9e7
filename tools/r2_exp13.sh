#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python tools/grad_steps.py block 4096 linear
timeout 300 python tools/grad_steps.py quadruped 8192
timeout 300 python tools/grad_steps.py ant 4096
DOJO_B200_LPT=2 timeout 300 python tools/grad_steps.py quadruped 8192
DOJO_B200_LPT=2 timeout 300 python tools/grad_steps.py block 4096 linear
DOJO_B200_NO_LS_ASSIST=1 timeout 300 python tools/grad_steps.py block 4096 linear
} > gpurun_out/r2_exp13.log 2>&1
cat gpurun_out/r2_exp13.log
