#!/bin/bash
B="python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --active-skip 0"
for v in q1 q8; do
  for m in static dynamic; do
    CLOVER_B200_QUEUE=$m CLOVER_B200_LIB=$PWD/cloverleaf_b200/libclover_b200_$v.so $B > gpurun_out/r2n_${v}_$m.json 2> gpurun_out/r2n_${v}_$m.err
  done
done
