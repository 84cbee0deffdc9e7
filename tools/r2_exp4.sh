#!/bin/bash
# round 2: ncu evidence for every benchmarked kernel + the second compilation; summaries are made on the box (reports are large)
mkdir -p gpurun_out
cap() {  # name skip count cmd...
  local name=$1 skip=$2 cnt=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:dojo_step_kernel -s $skip -c $cnt -f -o gpurun_out/prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  python tools/summarize_ncu.py gpurun_out/prof_$name.ncu-rep gpurun_out/r2_${name}_ncu_full > gpurun_out/r2_${name}_ncu.json 2>> gpurun_out/ncu_$name.log
}
{
echo "== quadruped parity test (stats in gpurun_out/parity_stats.jsonl)"
rm -f gpurun_out/parity_stats.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "step_parity or bench" 2>&1 | tail -15
echo "== second compilation (dj_cm): block with LinearContact / ImpactContact, sphere"
python tools/prof_one.py block 4096 5 fwd linear
python tools/prof_one.py block 4096 5 fwd impact
python tools/prof_one.py block 4096 5 fwd nonlinear
python tools/prof_one.py block 4096 5 grad linear
echo "== timing of the profiled configurations (no profiler)"
python tools/prof_one.py ant 4096 5 fwd
python tools/prof_one.py ant 4096 5 grad
python tools/prof_one.py quadruped 8192 5 fwd
python tools/prof_one.py quadruped 8192 5 grad
python tools/prof_one.py atlas 4096 3 fwd
python tools/prof_one.py ant 8192 5 fwd
python tools/prof_one.py ant 16384 5 fwd
echo "== launch list of bench.py (cold-cache, serialised: compare shares)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 3 --warmup 3 --no-sub > gpurun_out/bench_under_ncu.log 2>&1
tail -3 gpurun_out/bench_under_ncu.log | cut -c1-300
echo "== ncu --set full captures"
cap ant_fwd 23 1 python tools/prof_one.py ant 4096 1 fwd
cap ant_grad 23 2 python tools/prof_one.py ant 4096 1 grad
cap quadruped_fwd 4 1 python tools/prof_one.py quadruped 8192 1 fwd
cap atlas_fwd 4 1 python tools/prof_one.py atlas 4096 1 fwd
cap block_linear_cm 11 1 python tools/prof_one.py block 4096 1 fwd linear
ls -la gpurun_out/*.ncu-rep
# source-level hot spots of the forward kernel, made here so that only text travels back
ncu -i gpurun_out/prof_ant_fwd.ncu-rep --page source --csv > gpurun_out/r2_ant_fwd_source.csv 2>/dev/null
ls -la gpurun_out/r2_ant_fwd_source.csv
rm -f gpurun_out/prof_ant_grad.ncu-rep gpurun_out/prof_quadruped_fwd.ncu-rep gpurun_out/prof_atlas_fwd.ncu-rep gpurun_out/prof_block_linear_cm.ncu-rep
cat gpurun_out/r2_*_ncu.json
} > gpurun_out/r2_exp4.log 2>&1
tail -c 5000 gpurun_out/r2_exp4.log
