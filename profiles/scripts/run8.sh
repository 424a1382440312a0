#!/bin/bash
# 8-GPU box, final build: multi-GPU parity tests (verbose) + strong scaling of 3840^2 at N=1,2,4,8 (+trace at 8)
# + 15360^2 at N=8 and N=4 (the configuration gpurun flagged WEDGED in round 1)
TAG=${1:-x}
timeout 900 python -m pytest tests/test_dist.py -m gpu -v 2>&1 | grep -E "PASSED|FAILED|SKIPPED|passed|failed" > gpurun_out/${TAG}_dist8.log
run() { N=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 5 "$@"; }
run 8 --trace gpurun_out/${TAG}_t8 > gpurun_out/${TAG}_bench8.json 2> gpurun_out/${TAG}_bench8.err
run 4 > gpurun_out/${TAG}_bench4.json 2> gpurun_out/${TAG}_bench4.err
run 2 > gpurun_out/${TAG}_bench2.json 2> gpurun_out/${TAG}_bench2.err
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --active-skip 0 > gpurun_out/${TAG}_bench1.json 2> gpurun_out/${TAG}_bench1.err
run 8 --no-e2e --deck clover_bm256_short.in --steps 20 > gpurun_out/${TAG}_bench8_256.json 2> gpurun_out/${TAG}_bench8_256.err
run 4 --no-e2e --deck clover_bm256_short.in --steps 20 > gpurun_out/${TAG}_bench4_256.json 2> gpurun_out/${TAG}_bench4_256.err
nvidia-smi --query-gpu=index,memory.used,utilization.gpu --format=csv > gpurun_out/${TAG}_after.csv 2>&1
cat gpurun_out/${TAG}_dist8.log | tail -3
