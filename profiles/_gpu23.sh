#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t23.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t23b.log
timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab23.log 2>&1
AB_B=64 timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab23.log 2>&1
timeout 300 python profiles/explore_batch.py 8 16 32 64 > gpurun_out/r2_explore23.log 2>&1
tail -3 gpurun_out/r2_t23.log; tail -3 gpurun_out/r2_t23b.log; grep -v "^ \|Trace\|raise\|check" gpurun_out/r2_ab23.log; cat gpurun_out/r2_explore23.log
