#!/bin/bash
# GPU call D2 (2 GPUs): bench at --gpus 2 (async gathers, secondary configs under torchrun) + the NCCL test file.
set -u
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"; cut -c1-400 gpurun_out/bench_n2.json; grep -v "^\[W\|^W0" gpurun_out/bench_n2.err | tail -5
timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu -rs 2>&1 | tail -5 > gpurun_out/pytest_multi.log; cat gpurun_out/pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-200
