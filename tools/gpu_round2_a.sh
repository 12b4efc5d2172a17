#!/bin/bash
# GPU call A of round 2: the rest of the parity suite, compute-sanitizer over every kernel family, first numbers.
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
for tool in memcheck racecheck; do
  for fam in softmin ranges conv tc grid loss; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py $fam > gpurun_out/sanitize_${tool}_${fam}.log 2>&1
    echo "$tool $fam rc=$? $(grep -c 'ERROR SUMMARY' gpurun_out/sanitize_${tool}_${fam}.log) $(grep 'ERROR SUMMARY' gpurun_out/sanitize_${tool}_${fam}.log | tail -1)"
  done
done
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_n1.json
timeout 300 python tools/bench_configs.py multiscale 1000000 2>/dev/null | grep '^{' > gpurun_out/configs_ms.jsonl; cat gpurun_out/configs_ms.jsonl
timeout 300 python tools/bench_samplesloss.py 1000 10000 2>/dev/null | grep '^{' > gpurun_out/samplesloss.jsonl; cat gpurun_out/samplesloss.jsonl
for ns in 3 6 9 12 13 16 24 48; do timeout 120 ./build/explore 1000000 1000000 1e-4 3 "poly1/8" $ns | grep variant; done > gpurun_out/explore_nsplit.jsonl; cat gpurun_out/explore_nsplit.jsonl
for ns in 4 8 12 16; do timeout 120 ./build/explore 1000000 125000 1e-4 3 "poly1/8" $ns | grep variant; done > gpurun_out/explore_nsplit_shard.jsonl; cat gpurun_out/explore_nsplit_shard.jsonl
