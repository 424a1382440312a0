#!/bin/bash
# the driver's round-end sequence once more on the last commit (tests, smoke, default bench)
python -m pytest tests -m gpu -x -q 2>&1 | tail -n 4 > gpurun_out/r02_last_tests_1gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_last_smoke.log 2>&1
python bench.py > gpurun_out/r02_last_bench_default.json 2> gpurun_out/r02_last_bench_default.err
tail -n 2 gpurun_out/r02_last_tests_1gpu.log; tail -n 2 gpurun_out/r02_last_smoke.log
