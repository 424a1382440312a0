#!/bin/bash
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --active-skip 0 > gpurun_out/r3i_b1.json 2> gpurun_out/r3i_b1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r3i_b2.json 2> gpurun_out/r3i_b2.err
