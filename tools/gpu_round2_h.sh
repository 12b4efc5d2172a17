#!/bin/bash
set -u
mkdir -p gpurun_out
B200OT_LIB=$PWD/build/libb200ot_A.so python tools/ab_ops.py --tag A_ch8poly_unroll2 > gpurun_out/ab_A.jsonl 2> gpurun_out/ab_A.err; echo "A rc=$?"
python tools/ab_ops.py --tag B_ch16_unroll4 > gpurun_out/ab_B.jsonl 2> gpurun_out/ab_B.err; echo "B rc=$?"
B200OT_LIB=$PWD/build/libb200ot_C.so python tools/ab_ops.py --tag C_ch32_unroll8 > gpurun_out/ab_C.jsonl 2> gpurun_out/ab_C.err; echo "C rc=$?"
tail -3 gpurun_out/ab_*.err
