#!/bin/bash
# A/B on one box: the GPU tests on the in-tree build, then the same bench line for each variant library
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r3t_tests.log
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --active-skip 0"
$B > gpurun_out/r3t_main.json 2> gpurun_out/r3t_main.err
for v in tD pA xA; do
  CLOVER_B200_LIB=$PWD/cloverleaf_b200/libclover_b200_$v.so $B > gpurun_out/r3t_$v.json 2> gpurun_out/r3t_$v.err
done
$B > gpurun_out/r3t_main2.json 2> gpurun_out/r3t_main2.err
