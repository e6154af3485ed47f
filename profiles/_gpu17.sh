#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t17.log
timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab17.log 2>&1
AB_B=64 timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab17.log 2>&1
timeout 300 python profiles/explore_batch.py 8 64 > gpurun_out/r2_explore17.log 2>&1
ONLY_AUTO=1 timeout 300 python profiles/trace_epilogue.py > gpurun_out/r2_trace_epi17.log 2>&1
tail -3 gpurun_out/r2_t17.log; grep -v "^ \|Trace\|raise\|check" gpurun_out/r2_ab17.log; cat gpurun_out/r2_explore17.log; tail -30 gpurun_out/r2_trace_epi17.log
