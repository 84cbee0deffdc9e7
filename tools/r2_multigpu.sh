#!/bin/bash
# round 2, 8-GPU call: fused gather check with 8 ranks, bench --gpus 8 as the driver launches it
mkdir -p gpurun_out
{
nvidia-smi -L | wc -l
echo "== gather check, 8 ranks"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29731 tools/gather_check.py 2>&1 | grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" | tail -12
echo "== bench --gpus 8"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/bench_r2_n8.json 2> gpurun_out/bench_r2_n8.err
tail -c 400 gpurun_out/bench_r2_n8.err
python - <<'PY'
import json
for ln in open('gpurun_out/bench_r2_n8.json'):
    if not ln.startswith('{'): continue
    l = json.loads(ln)
    print({k: l[k] for k in ('value', 'ms_per_step', 'n_gpus', 'mean_newton_iters', 'failed_rate')}, 'e2e', l['e2e']['value'], 'gather', l.get('gather'))
    for k, r in l['sub_records'].items():
        print(k, {q: r.get(q) for q in ('value', 'ms_per_step', 'mean_newton_iters', 'failed_rate', 'error')}, 'e2e', r.get('e2e', {}).get('value'), r.get('gather'))
PY
} > gpurun_out/r2_multigpu.log 2>&1
tail -c 4000 gpurun_out/r2_multigpu.log
