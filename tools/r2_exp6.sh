#!/bin/bash
# round 2, call 2: compact body-body blocks + integer nanmax; full GPU suite; configs; LPT on / off; 2-GPU gather check is a separate call
mkdir -p gpurun_out
{
echo "== guarded first launch (TMA bulk staging of the plan tables is new)"
timeout 180 python tools/prof_one.py ant 256 2 fwd || { echo "FIRST LAUNCH FAILED"; exit 1; }
timeout 180 python tools/prof_one.py quadruped 256 2 grad
echo "== timing"
python tools/prof_one.py ant 4096 8 fwd
DOJO_B200_LPT=0 python tools/prof_one.py ant 4096 8 fwd
python tools/prof_one.py ant 4096 5 grad
python tools/prof_one.py quadruped 8192 5 fwd
python tools/prof_one.py quadruped 8192 5 grad
python tools/prof_one.py atlas 4096 3 fwd
python tools/prof_one.py atlas 1024 2 grad
python tools/prof_one.py ant 16384 5 fwd
DJ_PROF=1 DJ_ROLLOUT=0 python tools/time_variant.py build_variants/prof.so ant 4096 10
python tools/env_times.py build_variants/prof.so ant 4096 2>&1 | tail -8
echo "== gpu tests (whole suite)"
rm -f gpurun_out/parity_stats.jsonl
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -25
} > gpurun_out/r2_exp6.log 2>&1
tail -c 7000 gpurun_out/r2_exp6.log
