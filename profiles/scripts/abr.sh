#!/bin/bash
CLOVER_B200_XROW=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r3k_tests.log
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e"
$B > gpurun_out/r3k_main.json 2> gpurun_out/r3k_main.err
CLOVER_B200_XROW=1 $B > gpurun_out/r3k_xrow.json 2> gpurun_out/r3k_xrow.err
