#!/bin/bash
# GPU call U (last 3 GPU-minutes of the round): the tests that exercise what the measured defaults changed — merged GEMM-2
# instruction in the row-gradient kernels, one-pass value + gradient in every gaussian loss, forward routing — serially
set -u
mkdir -p gpurun_out
timeout 140 python -m pytest -q -m gpu --tb=line --no-header -p no:cacheprovider \
  tests/test_gpu_tc_variants.py tests/test_gpu_parity.py tests/test_gpu_reference_goldens.py \
  -k "one_pass or tensor_core_path or kernel_losses or kernel_conv or high_dimension or whole_losses or kernel_keops" \
  > gpurun_out/pytest_final_subset.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|Error" gpurun_out/pytest_final_subset.log | tail -30
