#!/bin/bash
# round 2, 2-GPU call: fused peer-write gather against NCCL (bit for bit), bench under torchrun (C4 / C3 shapes), single-GPU re-timing of the gradient path
mkdir -p gpurun_out
{
nvidia-smi -L
echo "== single GPU: gradient path after the fix of the backward sweep (overlapped / not overlapped)"
timeout 300 python tools/prof_one.py ant 4096 5 grad
DOJO_B200_NO_GRAD_OVERLAP=1 timeout 300 python tools/prof_one.py ant 4096 5 grad
timeout 300 python tools/prof_one.py ant 4096 8 fwd
timeout 300 python tools/prof_one.py quadruped 8192 5 grad
echo "== 2-GPU gather test"
timeout 600 python -m pytest tests/test_zz_gpu_gather.py tests/test_sharding.py -m gpu -q 2>&1 | tail -8
echo "== bench --gpus 2 (torchrun)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err
tail -c 600 gpurun_out/bench_r2_n2.err
python - <<'PY'
import json
l = json.load(open('gpurun_out/bench_r2_n2.json'))
print({k: l[k] for k in ('value', 'ms_per_step', 'n_gpus', 'mean_newton_iters', 'failed_rate')}, 'e2e', l['e2e']['value'], 'gather', l.get('gather'))
for k, r in l['sub_records'].items():
    print(k, {q: r.get(q) for q in ('value', 'ms_per_step', 'mean_newton_iters', 'failed_rate', 'error')}, 'e2e', r.get('e2e', {}).get('value'), r.get('gather'))
PY
echo "== bench --gpus 2 with the NCCL fallback gather"
DOJO_B200_GATHER=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 2 --steps 10 --warmup 3 --no-sub 2>/dev/null | cut -c1-400
} > gpurun_out/r2_exp7.log 2>&1
tail -c 6000 gpurun_out/r2_exp7.log
