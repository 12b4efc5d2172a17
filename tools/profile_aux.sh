#!/bin/bash
# Run on the GPU box (under gpurun): one full ncu capture of each secondary kernel AS SHIPPED — the tensor-core forward
# and row-gradient kernels at N = M = 1e6, D = 64 (BASELINE configs[2]), the separable grid pass at 256^3, the softmin
# row-gradient reduction at D = 3, the ranges-mode softmin of the multiscale fine phase and the small-problem iteration.
# tools/ncu_summary.py <rep> - r02 <name> turns each into profiles/r02_<name>_ncu.json.
set -u
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
$NCU -k regex:tc_reduce_kernel -s 1 -c 1 -o gpurun_out/prof_tc_fwd python tools/bench_conv.py 1000000 64 > gpurun_out/prof_tc_fwd.log 2>&1
$NCU -k regex:tc_bwd_kernel -s 1 -c 1 -o gpurun_out/prof_tc_bwd python tools/bench_conv.py 1000000 64 > gpurun_out/prof_tc_bwd.log 2>&1
$NCU -k regex:grid_pass_kernel -s 3 -c 1 -o gpurun_out/prof_grid_pass python tools/bench_configs.py grid 256 > gpurun_out/prof_grid.log 2>&1
$NCU -k regex:rowsum_partial_kernel -s 1 -c 1 -o gpurun_out/prof_rowsum python tools/bench_samplesloss.py 100000 > gpurun_out/prof_rowsum.log 2>&1
$NCU -k regex:softmin_partial_kernel.*true -s 40 -c 1 -o gpurun_out/prof_softmin_ranges python tools/bench_configs.py multiscale 1000000 > gpurun_out/prof_ranges.log 2>&1
$NCU -k regex:sinkhorn_iteration_small_kernel -s 5 -c 1 -o gpurun_out/prof_small_iter python tools/bench_samplesloss.py 1000 > gpurun_out/prof_small.log 2>&1
ls -la gpurun_out/*.ncu-rep
