#!/bin/bash
# 8-GPU box: multi-GPU parity tests + strong scaling of 3840^2 at N=8,4,2 (+trace) + 15360^2 at N=8
TAG=${1:-x}
timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/${TAG}_dist8.log
run() { N=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 5 "$@"; }
run 8 --trace gpurun_out/${TAG}_t8 > gpurun_out/${TAG}_bench8.json 2> gpurun_out/${TAG}_bench8.err
run 4 --no-e2e > gpurun_out/${TAG}_bench4.json 2> gpurun_out/${TAG}_bench4.err
run 2 --no-e2e > gpurun_out/${TAG}_bench2.json 2> gpurun_out/${TAG}_bench2.err
python bench.py --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --active-skip 0 > gpurun_out/${TAG}_bench1.json 2> gpurun_out/${TAG}_bench1.err
run 8 --no-e2e --deck clover_bm256_short.in --steps 20 > gpurun_out/${TAG}_bench8_256.json 2> gpurun_out/${TAG}_bench8_256.err
CLOVER_B200_XCTAS=64 run 8 --no-e2e > gpurun_out/${TAG}_bench8_x64.json 2> gpurun_out/${TAG}_bench8_x64.err
tail -2 gpurun_out/${TAG}_dist8.log
