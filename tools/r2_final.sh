#!/bin/bash
# round 2, first call of the second session: bench (both arms), tail analysis with the DJ_PROFILE build, ncu evidence, GPU tests
mkdir -p gpurun_out; rm -f gpurun_out/prof_*.ncu-rep
cap() {  # name skip count cmd...
  local name=$1 skip=$2 cnt=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:dojo_step_kernel -s $skip -c $cnt -f -o gpurun_out/prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1
  python tools/summarize_ncu.py gpurun_out/prof_$name.ncu-rep gpurun_out/r2_${name}_ncu_full > gpurun_out/r2_${name}_ncu.json 2>> gpurun_out/ncu_$name.log
}
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
echo "== bench (default line)"
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; tail -c 400 gpurun_out/bench_r2_final.err
python - <<'EOF'
import json
l = json.load(open('gpurun_out/bench_r2_final.json'))
print({k: l[k] for k in ('value', 'ms_per_step', 'mean_newton_iters', 'failed_rate', 'step_ms_min_max')}, l['e2e']['value'], l.get('rollout'), l.get('parity'), l.get('cpu_baseline', {}).get('value'))
for k, r in l['sub_records'].items():
    print(k, {q: r.get(q) for q in ('value', 'ms_per_step', 'mean_newton_iters', 'failed_rate', 'error')}, r.get('e2e', {}).get('value'), r.get('parity'))
EOF
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 | cut -c1-900
echo "== tail analysis (DJ_PROFILE build)"
python tools/env_times.py build_variants/prof.so ant 4096 2>&1 | tail -12
DJ_PROF=1 DJ_ROLLOUT=0 python tools/time_variant.py build_variants/prof.so ant 4096 10
echo "== timing of the profiled configurations (no profiler)"
python tools/prof_one.py ant 4096 5 fwd
python tools/prof_one.py ant 4096 5 grad
python tools/prof_one.py quadruped 8192 5 fwd
python tools/prof_one.py quadruped 8192 5 grad
python tools/prof_one.py atlas 4096 3 fwd
python tools/prof_one.py ant 16384 5 fwd
python tools/prof_one.py block 4096 5 fwd linear
python tools/prof_one.py block 4096 5 fwd impact
python tools/prof_one.py block 4096 5 fwd nonlinear
python tools/prof_one.py block 4096 5 grad linear
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
ncu -i gpurun_out/prof_ant_fwd.ncu-rep --page source --csv > gpurun_out/r2_ant_fwd_source.csv 2>/dev/null
ls -la gpurun_out/r2_ant_fwd_source.csv
rm -f gpurun_out/prof_ant_grad.ncu-rep gpurun_out/prof_quadruped_fwd.ncu-rep gpurun_out/prof_atlas_fwd.ncu-rep gpurun_out/prof_block_linear_cm.ncu-rep
cat gpurun_out/r2_*_ncu.json
echo "== gpu tests"
rm -f gpurun_out/parity_stats.jsonl
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25
} > gpurun_out/r2_final.log 2>&1
tail -c 6000 gpurun_out/r2_final.log
