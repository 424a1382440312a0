#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2z_tests.log
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r2z_main.json 2> gpurun_out/r2z_main.err
