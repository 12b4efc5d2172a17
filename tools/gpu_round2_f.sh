#!/bin/bash
# GPU call F: ring depth / tile width / off-load variants around the shipped softmin kernel, then the final N=1 bench.
set -u
mkdir -p gpurun_out
timeout 400 ./build/explore 1000000 1000000 1e-4 3 "r2 " 3 | grep variant > gpurun_out/explore_r2_variants.jsonl; cat gpurun_out/explore_r2_variants.jsonl | cut -c1-200
timeout 200 ./build/explore 1000000 125000 1e-4 3 "r2 " 3 | grep variant > gpurun_out/explore_r2_variants_shard.jsonl; cat gpurun_out/explore_r2_variants_shard.jsonl | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; cat gpurun_out/smoke.log | tail -6
timeout 700 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench1 rc=$?"; cut -c1-250 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
