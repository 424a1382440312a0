#!/bin/bash
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --active-skip 0"
$B > gpurun_out/r3h_main.json 2> gpurun_out/r3h_main.err
for v in q8 tt3 pt3 q2; do CLOVER_B200_LIB=$PWD/cloverleaf_b200/libclover_b200_$v.so $B > gpurun_out/r3h_$v.json 2> gpurun_out/r3h_$v.err; done
$B > gpurun_out/r3h_main2.json 2> gpurun_out/r3h_main2.err
