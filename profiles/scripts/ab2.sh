#!/bin/bash
timeout 600 python -m pytest tests/test_dist.py -m gpu -v 2>&1 | grep -E "PASSED|FAILED|SKIPPED|passed|failed" > gpurun_out/r02_final_dist2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 > gpurun_out/r02_final2_bench2.json 2> gpurun_out/r02_final2_bench2.err
tail -3 gpurun_out/r02_final_dist2.log
