#!/bin/bash
# 1 GPU: launch list of the bench command + ncu --set full of the seven production kernels (steady-state step)
TAG=${1:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --active-skip 0 --profile-steps 0 > gpurun_out/${TAG}_launches_bench.json 2> gpurun_out/${TAG}_launches_bench.err
ncu --set full --clock-control none --import-source on -k regex:'tma_kernel' -s 56 -c 7 -o gpurun_out/${TAG}_full -f \
    python profiles/ncu_step.py 12 > gpurun_out/${TAG}_full.log 2>&1
ls -la gpurun_out/${TAG}_full.ncu-rep
