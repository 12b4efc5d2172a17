#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ./build/explore 1000000 1000000 1e-4 3 "r3 " 3 | grep variant > gpurun_out/explore_r3_variants.jsonl; cut -c1-190 gpurun_out/explore_r3_variants.jsonl
timeout 300 ./build/explore 1000000 125000 1e-4 3 "r3 " 3 | grep variant > gpurun_out/explore_r3_variants_shard.jsonl; cut -c1-190 gpurun_out/explore_r3_variants_shard.jsonl
