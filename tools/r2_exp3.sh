#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_stats.jsonl
{
echo "== gpu tests"
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -40
echo "== bench v2"
timeout 900 python bench.py --steps 10 --warmup 3
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 5 --warmup 3
} > gpurun_out/r2_exp3.log 2>&1
tail -c 6000 gpurun_out/r2_exp3.log
