#!/bin/bash
# compute-sanitizer on a small ragged mesh (every TMA kernel runs, tiles cut at both edges), final build
timeout 70 compute-sanitizer --tool racecheck python profiles/ncu_step.py 3 clover_bm_short.in 250 130 > gpurun_out/r02_final_san_racecheck.log 2>&1
timeout 50 compute-sanitizer --tool memcheck python profiles/ncu_step.py 3 clover_bm_short.in 250 130 > gpurun_out/r02_final_san_memcheck.log 2>&1
tail -n 2 gpurun_out/r02_final_san_racecheck.log; tail -n 2 gpurun_out/r02_final_san_memcheck.log
