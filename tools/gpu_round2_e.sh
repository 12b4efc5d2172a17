#!/bin/bash
# GPU call E: racecheck of the TMA ring with per-thread stage release (dense with forced reuse, ranges, loss), headline
# kernel timing after the change, quick parity re-check.
set -u
mkdir -p gpurun_out
for ns in 3 3; do timeout 120 ./build/explore 1000000 1000000 1e-4 3 "guard poly1/8" $ns | grep variant; done > gpurun_out/explore_after_release.jsonl; cat gpurun_out/explore_after_release.jsonl
B200OT_FORCE_SPLITS=1 timeout 600 compute-sanitizer --tool racecheck --print-limit 6 python tools/sanitize_smoke.py ring > gpurun_out/sanitize_racecheck_ring.log 2>&1; echo "ring: $(grep 'RACECHECK SUMMARY' gpurun_out/sanitize_racecheck_ring.log | tail -1)"
for fam in ranges loss; do
  timeout 600 compute-sanitizer --tool racecheck --print-limit 6 python tools/sanitize_smoke.py $fam > gpurun_out/sanitize_racecheck_${fam}.log 2>&1
  echo "racecheck $fam: $(grep 'RACECHECK SUMMARY' gpurun_out/sanitize_racecheck_${fam}.log | tail -1)"
done
B200OT_FORCE_SPLITS=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 6 python tools/sanitize_smoke.py ring > gpurun_out/sanitize_memcheck_ring.log 2>&1; echo "ring memcheck: $(grep 'ERROR SUMMARY' gpurun_out/sanitize_memcheck_ring.log | tail -1)"
timeout 900 python -m pytest tests -q -m gpu -x -k "softmin or sinkhorn_cases or keops or full_size or multiscale or kernel" 2>&1 | tail -5
