#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t21.log
timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab21.log 2>&1
AB_B=64 timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab21.log 2>&1
for w in 2 8; do AB_B=64 LDM_B200_T_GN_WANT=$w timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab21.log 2>&1; done
timeout 300 python profiles/explore_batch.py 8 64 > gpurun_out/r2_explore21.log 2>&1
ONLY_AUTO=1 timeout 300 python profiles/trace_epilogue.py > gpurun_out/r2_trace_epi21.log 2>&1
python profiles/one_gn.py > gpurun_out/r2_gn21.log 2>&1
tail -3 gpurun_out/r2_t21.log; grep -v "^ \|Trace\|raise\|check" gpurun_out/r2_ab21.log; cat gpurun_out/r2_explore21.log; tail -2 gpurun_out/r2_gn21.log; grep "lean\|FF2\|GEGLU (row" gpurun_out/r2_trace_epi21.log | cut -c1-250
