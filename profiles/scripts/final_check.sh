#!/bin/bash
# what the driver does at round end, on the final commit: GPU tests, smoke(), default bench; + ncu of the two march kernels
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02_final_tests_1gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1
python bench.py > gpurun_out/r02_final_bench_default.json 2> gpurun_out/r02_final_bench_default.err
ncu --set full --clock-control none --import-source on -k regex:'ymarch' -s 8 -c 2 -o gpurun_out/r02_march_full -f \
    python profiles/ncu_step.py 6 > gpurun_out/r02_march_full.log 2>&1
tail -2 gpurun_out/r02_final_tests_1gpu.log; cat gpurun_out/r02_final_smoke.log | tail -2
