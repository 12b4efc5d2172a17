#!/bin/bash
# GPU call O: warps per CTA / tile width of the small-problem kernels (A/B builds)
set -u
mkdir -p gpurun_out
: > gpurun_out/ab_small.jsonl
python tools/ab_small.py --tag W4T256 >> gpurun_out/ab_small.jsonl 2> gpurun_out/ab_small.err
for v in W8T256 W16T256 W8T512 W16T512; do
  B200OT_LIB=$PWD/build/libb200ot_$v.so python tools/ab_small.py --tag $v >> gpurun_out/ab_small.jsonl 2>> gpurun_out/ab_small.err
done
tail -3 gpurun_out/ab_small.err; wc -l gpurun_out/ab_small.jsonl
