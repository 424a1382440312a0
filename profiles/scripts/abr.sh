#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r3a_tests.log
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --active-skip 0 --trace gpurun_out/r3a_t1 > gpurun_out/r3a_main.json 2> gpurun_out/r3a_main.err
