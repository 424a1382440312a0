#!/bin/bash
# GPU tests + a short bench line (no CPU baseline, no e2e) on the current tree
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3 > gpurun_out/r02_quick_tests.log
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --active-skip 0 > gpurun_out/r02_quick_bench.json 2> gpurun_out/r02_quick_bench.err
tail -n 1 gpurun_out/r02_quick_tests.log
