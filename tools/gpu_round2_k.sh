#!/bin/bash
# GPU call K: shortened last tile — kernel timing (full problem, 8-GPU shard), parity
set -u
mkdir -p gpurun_out
for M in 1000000 125000; do
  timeout 300 ./build/explore 1000000 $M 1e-4 3 "r3 ch16 nopoly" | grep variant | sed "s/r3 ch16 nopoly/prepass256+tail ch16 nopoly/"
done > gpurun_out/explore_tail.jsonl; cut -c1-200 gpurun_out/explore_tail.jsonl
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_goldens.py -q -m gpu -x 2>&1 | tail -4
