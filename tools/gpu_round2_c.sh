#!/bin/bash
# GPU call C (2 GPUs): the complete GPU suite (final log), NCCL tests on 2 ranks, bench at --gpus 1 and 2.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench1 rc=$?"; cut -c1-300 gpurun_out/bench_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"; cut -c1-300 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
timeout 200 python tools/bench_configs.py grid 256 2>/dev/null | grep '^{' > gpurun_out/configs_grid.jsonl; cat gpurun_out/configs_grid.jsonl
