#!/bin/bash
# GPU call P: full validation of the final tree — GPU suite, smoke, bench, sanitizer on the families touched last
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; cat gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench1 rc=$?"; cut -c1-300 gpurun_out/bench_n1.json
python tools/bench_samplesloss.py 1000 5000 > gpurun_out/samplesloss_final.jsonl 2>/dev/null; cat gpurun_out/samplesloss_final.jsonl
python tools/profile_small_host.py 1000 sinkhorn 2>&1 | head -1
for fam in loss softmin; do
  timeout 600 compute-sanitizer --tool racecheck --print-limit 6 python tools/sanitize_smoke.py $fam > gpurun_out/sanitize_racecheck_${fam}.log 2>&1
  echo "racecheck $fam: $(grep 'RACECHECK SUMMARY' gpurun_out/sanitize_racecheck_${fam}.log | tail -1)"
  timeout 600 compute-sanitizer --tool memcheck --print-limit 6 python tools/sanitize_smoke.py $fam > gpurun_out/sanitize_memcheck_${fam}.log 2>&1
  echo "memcheck $fam: $(grep 'ERROR SUMMARY' gpurun_out/sanitize_memcheck_${fam}.log | tail -1)"
done
timeout 600 compute-sanitizer --tool memcheck --print-limit 6 python tools/sanitize_smoke.py ranges > gpurun_out/sanitize_memcheck_ranges.log 2>&1; echo "memcheck ranges: $(grep 'ERROR SUMMARY' gpurun_out/sanitize_memcheck_ranges.log | tail -1)"
