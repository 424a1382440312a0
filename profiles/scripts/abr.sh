#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r3p_tests.log
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --active-skip 0"
$B > gpurun_out/r3p_main.json 2> gpurun_out/r3p_main.err
