#!/bin/bash
# GPU call L (8 GPUs): bench at --gpus 8 (secondary configs incl. cfg4 N=1e7) + the NCCL parity tests (world 4)
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus 8 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench8 rc=$?"; cut -c1-400 gpurun_out/bench_n8.json; grep -v "^\[W\|^W0" gpurun_out/bench_n8.err | tail -5
timeout 400 python -m pytest tests/test_gpu_multi.py -q -m gpu -rs 2>&1 | tail -5 > gpurun_out/pytest_multi.log; cat gpurun_out/pytest_multi.log
