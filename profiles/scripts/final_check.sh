#!/bin/bash
# what the driver does at round end, on the final commit: GPU tests, smoke(), default bench; then the ncu launch list of
# the bench command and ncu --set full of one steady-state step (seven production kernels)
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02_final_tests_1gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1
python bench.py > gpurun_out/r02_final_bench_default.json 2> gpurun_out/r02_final_bench_default.err
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --active-skip 0 --profile-steps 0 > gpurun_out/r02_final_launches_bench.json 2> gpurun_out/r02_final_launches_bench.err
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'tma_kernel' -s 56 -c 7 -o gpurun_out/r02_final_full -f \
    python profiles/ncu_step.py 12 > gpurun_out/r02_final_full.log 2>&1
tail -2 gpurun_out/r02_final_tests_1gpu.log; tail -2 gpurun_out/r02_final_smoke.log
