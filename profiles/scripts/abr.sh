#!/bin/bash
CLOVER_B200_MOM_YMARCH=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r3l_tests.log
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e"
$B > gpurun_out/r3l_main.json 2> gpurun_out/r3l_main.err
CLOVER_B200_MOM_YMARCH=1 $B > gpurun_out/r3l_march.json 2> gpurun_out/r3l_march.err
