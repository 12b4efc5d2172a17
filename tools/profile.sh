#!/bin/bash
# Run on the GPU box (under gpurun): launch list + one full ncu capture of the dominant kernel for the
# exact bench command.  Outputs land in gpurun_out/; tools/ncu_summary.py turns them into profiles/*.
set -u
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary"
# 1. every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    $BENCH > gpurun_out/launches_bench.log 2>&1
# 2. the dominant kernel, full set, source-level (skip the warm-up launches)
ncu --set full --clock-control none --import-source on -k regex:softmin_partial_kernel -s 8 -c 1 \
    -o gpurun_out/prof_softmin_partial -f $BENCH > gpurun_out/prof_bench.log 2>&1
ls -la gpurun_out/
