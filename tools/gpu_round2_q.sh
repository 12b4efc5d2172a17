#!/bin/bash
# GPU call Q: where does the time of the two-scale Sinkhorn (configs[3], N=1e6) go?  Launch list under ncu.
set -u
mkdir -p gpurun_out
python tools/run_multiscale_once.py 1000000
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_multiscale.csv python tools/run_multiscale_once.py 1000000 > gpurun_out/launches_multiscale.log 2>&1
tail -2 gpurun_out/launches_multiscale.log; wc -l gpurun_out/launches_multiscale.csv
