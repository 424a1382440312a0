#!/bin/bash
TAG=${1:-x}
timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${TAG}_dist4.log
run() { N=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 40 --warmup 5 --no-e2e "$@"; }
run 4 --trace gpurun_out/${TAG}_t4 > gpurun_out/${TAG}_b4.json 2> gpurun_out/${TAG}_b4.err
run 2 > gpurun_out/${TAG}_b2.json 2> gpurun_out/${TAG}_b2.err
CLOVER_B200_XCTAS=64 run 4 > gpurun_out/${TAG}_b4_x64.json 2> gpurun_out/${TAG}_b4_x64.err
CLOVER_B200_XCTAS=16 run 4 > gpurun_out/${TAG}_b4_x16.json 2> gpurun_out/${TAG}_b4_x16.err
tail -2 gpurun_out/${TAG}_dist4.log
