#!/bin/bash
# GPU call B of round 2: debug the p = 1 small-path mismatch, nsplit 1..3, racecheck of the ranges mode, small-N timing,
# ncu captures of every kernel family as shipped.
set -u
mkdir -p gpurun_out
timeout 300 python tools/debug_p1.py > gpurun_out/debug_p1.log 2>&1; cat gpurun_out/debug_p1.log | tail -12
timeout 900 python -m pytest tests -q -m gpu -x --tb=short -k "multiscale_vs_two_scale or keops_backends or live_reference or high_dimension or full_sinkhorn_loop or grid_softmin_256 or tensor_core" 2>&1 | tail -60 > gpurun_out/pytest_focus.log; tail -25 gpurun_out/pytest_focus.log
for ns in 1 2 3; do timeout 120 ./build/explore 1000000 1000000 1e-4 3 "guard poly1/8" $ns | grep variant; done > gpurun_out/explore_nsplit_low.jsonl; cat gpurun_out/explore_nsplit_low.jsonl
for ns in 1 2 3; do timeout 120 ./build/explore 1000000 125000 1e-4 3 "guard poly1/8" $ns | grep variant; done > gpurun_out/explore_nsplit_shard_low.jsonl; cat gpurun_out/explore_nsplit_shard_low.jsonl
for fam in ranges loss; do
  timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_smoke.py $fam > gpurun_out/sanitize_racecheck_${fam}.log 2>&1
  echo "racecheck $fam rc=$? $(grep 'RACECHECK SUMMARY' gpurun_out/sanitize_racecheck_${fam}.log | tail -1)"
done
timeout 300 python tools/bench_samplesloss.py 1000 5000 2>/dev/null | grep '^{' > gpurun_out/samplesloss.jsonl; cat gpurun_out/samplesloss.jsonl
timeout 900 bash tools/profile.sh > gpurun_out/profile_sh.log 2>&1
timeout 1200 bash tools/profile_aux.sh > gpurun_out/profile_aux.log 2>&1; tail -8 gpurun_out/profile_aux.log
