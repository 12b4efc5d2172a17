#!/bin/bash
# GPU call N: small-path host overhead — fused value kernels, one-launch bounding box, cached uniform weights
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_goldens.py tests/test_live_reference.py -q -m gpu -x 2>&1 | tail -6
python tools/profile_small_host.py 1000 sinkhorn > gpurun_out/profile_small_host_after.txt 2>&1; head -3 gpurun_out/profile_small_host_after.txt
python tools/profile_small_host.py 1000 gaussian > gpurun_out/profile_small_host_gauss.txt 2>&1; head -3 gpurun_out/profile_small_host_gauss.txt
python tools/bench_samplesloss.py 1000 5000 > gpurun_out/samplesloss_after.jsonl 2>&1; cat gpurun_out/samplesloss_after.jsonl
