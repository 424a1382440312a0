#!/bin/bash
# 1-GPU full tests + 1- and 2-GPU bench with in-situ traces and exchange-grid A/B (run under gpurun --gpus 2)
TAG=${1:-x}
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/${TAG}_tests.log
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --trace gpurun_out/${TAG}_t1 > gpurun_out/${TAG}_bench1.json 2> gpurun_out/${TAG}_bench1.err
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 --no-e2e"
timeout 300 $B --trace gpurun_out/${TAG}_t2 > gpurun_out/${TAG}_bench2.json 2> gpurun_out/${TAG}_bench2.err
CLOVER_B200_XCTAS=148 timeout 300 $B > gpurun_out/${TAG}_bench2_x148.json 2> gpurun_out/${TAG}_bench2_x148.err
CLOVER_B200_XCTAS=8 timeout 300 $B > gpurun_out/${TAG}_bench2_x8.json 2> gpurun_out/${TAG}_bench2_x8.err
tail -3 gpurun_out/${TAG}_tests.log
