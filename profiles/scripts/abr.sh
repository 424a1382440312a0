#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r3g_tests.log
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --active-skip 0"
$B > gpurun_out/r3g_fast.json 2> gpurun_out/r3g_fast.err
CLOVER_B200_LIB=$PWD/cloverleaf_b200/libclover_b200_lcslow.so $B > gpurun_out/r3g_slow.json 2> gpurun_out/r3g_slow.err
$B > gpurun_out/r3g_fast2.json 2> gpurun_out/r3g_fast2.err
