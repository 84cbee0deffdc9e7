#!/bin/bash
# round 2: line-search assist by the drained slots of a CTA (tail of the launch) on / off
mkdir -p gpurun_out
{
echo "== guarded first launches (new CTA-wide protocol)"
timeout 120 python tools/prof_one.py ant 37 2 fwd || { echo "FIRST LAUNCH FAILED"; exit 1; }
timeout 120 python tools/prof_one.py atlas 5 2 fwd || { echo "ATLAS LAUNCH FAILED"; exit 1; }
echo "== assist on / off: time and checksum (must be identical)"
DJ_ROLLOUT=0 timeout 300 python tools/time_variant.py dojo.jl_b200/libdojo_b200.so ant 4096 10 | cut -c1-200
DOJO_B200_NO_LS_ASSIST=1 DJ_ROLLOUT=0 timeout 300 python tools/time_variant.py dojo.jl_b200/libdojo_b200.so ant 4096 10 | cut -c1-200
timeout 300 python tools/prof_one.py ant 4096 8 fwd
DOJO_B200_NO_LS_ASSIST=1 timeout 300 python tools/prof_one.py ant 4096 8 fwd
timeout 300 python tools/prof_one.py ant 4096 5 grad
timeout 300 python tools/prof_one.py quadruped 8192 5 fwd
DOJO_B200_NO_LS_ASSIST=1 timeout 300 python tools/prof_one.py quadruped 8192 5 fwd
timeout 300 python tools/prof_one.py quadruped 8192 5 grad
timeout 300 python tools/prof_one.py atlas 4096 3 fwd
DOJO_B200_NO_LS_ASSIST=1 timeout 300 python tools/prof_one.py atlas 4096 3 fwd
timeout 300 python tools/prof_one.py ant 1024 8 fwd
DOJO_B200_NO_LS_ASSIST=1 timeout 300 python tools/prof_one.py ant 1024 8 fwd
echo "== gpu tests (whole suite)"
rm -f gpurun_out/parity_stats.jsonl
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -8
} > gpurun_out/r2_exp10.log 2>&1
grep -v "config:" gpurun_out/r2_exp10.log | tail -c 5000
