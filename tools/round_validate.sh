#!/bin/bash
# One gpurun call that refreshes every measured artefact of a round (≈ 15 minutes of box time on one B200):
#   gpurun --timeout 2400 -- 'bash tools/round_validate.sh r03'
# then, back in the build container:
#   python tools/ncu_summary.py gpurun_out/prof_softmin_partial.ncu-rep gpurun_out/launches.csv r03
#   for n in tc_fwd tc_bwd grid_pass rowsum small_iter; do python tools/ncu_summary.py gpurun_out/prof_$n.ncu-rep - r03 $n; done
#   python tools/sass_summary.py r03
# and copy gpurun_out/{pytest_gpu.log,bench_n1.json,samplesloss.jsonl,sanitize_*.log} into profiles/<tag>_*.
set -u
TAG=${1:-rXX}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 1500 python -m pytest tests -q -m gpu -rs > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for tool in memcheck racecheck; do
  for fam in softmin ranges conv tc grid loss; do
    timeout 400 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py $fam > gpurun_out/sanitize_${tool}_${fam}.log 2>&1
    echo "$tool $fam: $(grep 'SUMMARY' gpurun_out/sanitize_${tool}_${fam}.log | tail -1)"
  done
done
timeout 700 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 300 python tools/ab_tc_route.py > gpurun_out/ab_tc_route.jsonl 2>/dev/null
timeout 300 python tools/bench_samplesloss.py 1000 5000 10000 100000 2>/dev/null | grep '^{' > gpurun_out/samplesloss.jsonl
timeout 600 bash tools/profile.sh > gpurun_out/profile_sh.log 2>&1
timeout 1200 bash tools/profile_aux.sh > gpurun_out/profile_aux.log 2>&1
echo "tag=$TAG"; ls gpurun_out | tr '\n' ' '
