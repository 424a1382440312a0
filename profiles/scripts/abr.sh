#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2u_tests.log
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --active-skip 0"
$B --trace gpurun_out/r2u_t1 > gpurun_out/r2u_main.json 2> gpurun_out/r2u_main.err
CLOVER_B200_MERGE_HALO=0 $B --no-e2e > gpurun_out/r2u_nomerge.json 2> gpurun_out/r2u_nomerge.err
CLOVER_B200_LIB=$PWD/cloverleaf_b200/libclover_b200_r2c.so $B --no-e2e > gpurun_out/r2u_r2c.json 2> gpurun_out/r2u_r2c.err
