#!/bin/bash
# GPU call D1 (1 GPU): complete suite, bench N=1, ncu of the final softmin plan + ranges kernel, racecheck detail.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -rs 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
timeout 700 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench1 rc=$?"; cut -c1-300 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 600 bash tools/profile.sh > gpurun_out/profile_sh.log 2>&1
ncu --set full --clock-control none --import-source on -f -k regex:softmin_partial_kernel -s 60 -c 1 -o gpurun_out/prof_softmin_ranges python tools/bench_configs.py multiscale 1000000 > gpurun_out/prof_ranges.log 2>&1
timeout 400 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 6 python tools/sanitize_smoke.py ranges > gpurun_out/sanitize_racecheck_ranges_all.log 2>&1; grep -c "hazard" gpurun_out/sanitize_racecheck_ranges_all.log
timeout 200 python tools/bench_samplesloss.py 1000 2>/dev/null | grep '^{' > gpurun_out/samplesloss_1k.jsonl; cat gpurun_out/samplesloss_1k.jsonl
