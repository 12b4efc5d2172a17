#!/bin/bash
# GPU call J: max pre-pass — kernel timing on the full problem and on the 8-GPU shard, parity subset
set -u
mkdir -p gpurun_out
for M in 1000000 125000; do
  timeout 300 ./build/explore 1000000 $M 1e-4 3 "r3 ch16 nopoly" | grep variant | sed "s/r3 ch16 nopoly/prepass256 ch16 nopoly/"
done > gpurun_out/explore_prepass.jsonl; cut -c1-200 gpurun_out/explore_prepass.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_goldens.py -q -m gpu -x 2>&1 | tail -4
timeout 600 python tools/ab_ops.py --tag prepass > gpurun_out/ab_prepass.jsonl 2> gpurun_out/ab_prepass.err; echo "ab rc=$?"
