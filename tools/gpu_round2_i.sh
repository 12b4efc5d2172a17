#!/bin/bash
# GPU call I: validation of the per-(D,p) chunk lengths — seeded-max experiment, full GPU suite, smoke, bench, ncu, racecheck
set -u
mkdir -p gpurun_out
for bin in explore explore_seed; do
  for M in 1000000 125000; do
    timeout 300 ./build/$bin 1000000 $M 1e-4 3 "r3 ch16 nopoly" | grep variant | sed "s/r3 ch16 nopoly/$bin ch16 nopoly/"
  done
done > gpurun_out/explore_seed.jsonl; cut -c1-200 gpurun_out/explore_seed.jsonl
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; cat gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench1 rc=$?"; cut -c1-400 gpurun_out/bench_n1.json
timeout 900 bash tools/profile.sh > /dev/null 2>&1; echo "profile rc=$?"
for fam in softmin ranges; do
  timeout 600 compute-sanitizer --tool racecheck --print-limit 6 python tools/sanitize_smoke.py $fam > gpurun_out/sanitize_racecheck_${fam}.log 2>&1
  echo "racecheck $fam: $(grep 'RACECHECK SUMMARY' gpurun_out/sanitize_racecheck_${fam}.log | tail -1)"
done
B200OT_FORCE_SPLITS=1 timeout 600 compute-sanitizer --tool racecheck --print-limit 6 python tools/sanitize_smoke.py ring > gpurun_out/sanitize_racecheck_ring.log 2>&1; echo "ring: $(grep 'RACECHECK SUMMARY' gpurun_out/sanitize_racecheck_ring.log | tail -1)"
timeout 600 compute-sanitizer --tool memcheck --print-limit 6 python tools/sanitize_smoke.py softmin > gpurun_out/sanitize_memcheck_softmin.log 2>&1; echo "memcheck softmin: $(grep 'ERROR SUMMARY' gpurun_out/sanitize_memcheck_softmin.log | tail -1)"
