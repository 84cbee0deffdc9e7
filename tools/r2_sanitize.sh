#!/bin/bash
# compute-sanitizer over a small invocation of every C-ABI entry point (tools/sanity.py): memcheck, racecheck (shared-memory hazards:
# the CTA-wide line-search mailbox, the pair shuffles of the joint assembly), synccheck
mkdir -p gpurun_out
{
for tool in memcheck racecheck synccheck; do
  for m in ant quadruped; do
    echo "== compute-sanitizer --tool $tool  tools/sanity.py $m"
    timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanity.py $m 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanity ok|widened ok|Error|error|hazard" | head -12
  done
done
echo "== memcheck, second compilation"
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanity.py block:linear 2>&1 | grep -E "ERROR SUMMARY|sanity ok|Error" | head -5
} > gpurun_out/r2_sanitize.log 2>&1
cat gpurun_out/r2_sanitize.log
