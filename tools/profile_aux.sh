#!/bin/bash
# Run on the GPU box (under gpurun): one full ncu capture of each secondary kernel (tensor-core forward and
# backward at D = 64, the separable grid pass at 256^3, the softmin row-gradient reduction at D = 3).
# tools/ncu_summary.py <rep> - r01 <name> turns each into profiles/r01_<name>_ncu.json.
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:tc_reduce_kernel -s 1 -c 1 -o gpurun_out/prof_tc_fwd python tools/bench_conv.py 200000 64 > gpurun_out/prof_tc_fwd.log 2>&1
$NCU -k regex:tc_bwd_kernel -s 1 -c 1 -o gpurun_out/prof_tc_bwd python tools/bench_conv.py 200000 64 > gpurun_out/prof_tc_bwd.log 2>&1
$NCU -k regex:grid_pass_kernel -s 3 -c 1 -o gpurun_out/prof_grid_pass python tools/bench_configs.py grid 256 > gpurun_out/prof_grid.log 2>&1
$NCU -k regex:rowsum_partial_kernel -s 1 -c 1 -o gpurun_out/prof_rowsum python tools/bench_samplesloss.py 100000 > gpurun_out/prof_rowsum.log 2>&1
ls -la gpurun_out/*.ncu-rep
