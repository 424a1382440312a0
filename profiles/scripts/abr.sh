#!/bin/bash
# current build vs the r2c-state build on the same box (+ full GPU tests)
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2m_tests.log
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --active-skip 0"
CLOVER_B200_LIB=$PWD/cloverleaf_b200/libclover_b200_r2c.so $B > gpurun_out/r2m_r2c.json 2> gpurun_out/r2m_r2c.err
$B > gpurun_out/r2m_main.json 2> gpurun_out/r2m_main.err
CLOVER_B200_QUEUE=static $B > gpurun_out/r2m_static.json 2> gpurun_out/r2m_static.err
CLOVER_B200_PDL=0 $B > gpurun_out/r2m_nopdl.json 2> gpurun_out/r2m_nopdl.err
