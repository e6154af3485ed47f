#!/bin/bash
mkdir -p gpurun_out
profiles/ubench/f32x2 > gpurun_out/r2_ubench_f32x2.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t18.log
timeout 300 python profiles/trace_attn.py > gpurun_out/r2_trace_attn18.log 2>&1
timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab18.log 2>&1
AB_B=64 timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab18.log 2>&1
timeout 300 python profiles/explore_batch.py 8 64 > gpurun_out/r2_explore18.log 2>&1
cat gpurun_out/r2_ubench_f32x2.txt; tail -3 gpurun_out/r2_t18.log; grep -v "^ \|Trace\|raise\|check" gpurun_out/r2_ab18.log; cat gpurun_out/r2_explore18.log; grep -A3 "^n=16 t=1024 tk=1024" gpurun_out/r2_trace_attn18.log;  grep "^n=" gpurun_out/r2_trace_attn18.log
