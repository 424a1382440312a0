#!/bin/bash
TAG=${1:-x}
timeout 600 python -m pytest tests/test_dist.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/${TAG}_dist.log
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5"
timeout 300 $B --trace gpurun_out/${TAG}_t2 > gpurun_out/${TAG}_bench2.json 2> gpurun_out/${TAG}_bench2.err
tail -2 gpurun_out/${TAG}_dist.log
