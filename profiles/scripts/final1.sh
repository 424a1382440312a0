#!/bin/bash
# 1 GPU: full bench line (CPU baseline + active regime + e2e), 15360^2 on one GPU, compute-sanitizer on a small mesh
TAG=${1:-r02}
python bench.py --steps 40 --warmup 5 > gpurun_out/${TAG}_bench_full.json 2> gpurun_out/${TAG}_bench_full.err
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --active-skip 0 --deck clover_bm256_short.in > gpurun_out/${TAG}_bench1_256.json 2> gpurun_out/${TAG}_bench1_256.err
python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --active-skip 0 --deck clover_bm64_short.in > gpurun_out/${TAG}_bench1_64.json 2> gpurun_out/${TAG}_bench1_64.err
timeout 400 compute-sanitizer --tool memcheck python profiles/ncu_step.py 3 clover_bm_short.in 250 130 > gpurun_out/${TAG}_san_memcheck.log 2>&1
timeout 400 compute-sanitizer --tool racecheck python profiles/ncu_step.py 3 clover_bm_short.in 250 130 > gpurun_out/${TAG}_san_racecheck.log 2>&1
tail -3 gpurun_out/${TAG}_san_memcheck.log gpurun_out/${TAG}_san_racecheck.log
