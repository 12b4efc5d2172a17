#!/bin/bash
# One gpurun call that refreshes every measured artefact of a round (≈ 8 minutes of box time on one B200):
#   gpurun --timeout 1500 -- 'bash tools/round_validate.sh r02'
# then, back in the build container:
#   python tools/ncu_summary.py gpurun_out/prof_softmin_partial.ncu-rep gpurun_out/launches.csv r02
#   for n in tc_fwd tc_bwd grid_pass rowsum; do python tools/ncu_summary.py gpurun_out/prof_$n.ncu-rep - r02 $n; done
# and copy gpurun_out/{pytest_gpu.log,bench_n1.json,conv_bench.jsonl,configs.jsonl,samplesloss.jsonl} into profiles/<tag>_*.
set -u
TAG=${1:-rXX}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 200 python tools/bench_conv.py 400000 16 32 64 2>/dev/null | grep '^{' > gpurun_out/conv_bench.jsonl
timeout 300 python tools/bench_configs.py grid 128 256 2>/dev/null | grep '^{' > gpurun_out/configs.jsonl
timeout 300 python tools/bench_configs.py multiscale 1000000 2>/dev/null | grep '^{' >> gpurun_out/configs.jsonl
timeout 300 python tools/bench_samplesloss.py 1000 10000 100000 2>/dev/null | grep '^{' > gpurun_out/samplesloss.jsonl
timeout 600 bash tools/profile.sh > gpurun_out/profile_sh.log 2>&1
timeout 600 bash tools/profile_aux.sh > gpurun_out/profile_aux.log 2>&1
echo "tag=$TAG"; ls gpurun_out | tr '\n' ' '
