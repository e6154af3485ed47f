#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t16.log
for n in 16 128; do sed "s/for (n, t, tk, heads, d) in \[(16,/for (n, t, tk, heads, d) in [($n,/" profiles/trace_attn.py > /dev/null; done
timeout 300 python profiles/trace_attn.py > gpurun_out/r2_trace_attn16.log 2>&1
for sw in "X=1" "LDM_B200_POLY_EXP=0"; do env $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab16.log 2>&1; done
for sw in "X=1" "LDM_B200_POLY_EXP=0"; do env AB_B=64 $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab16.log 2>&1; done
timeout 300 python profiles/explore_batch.py 8 64 > gpurun_out/r2_explore16.log 2>&1
tail -3 gpurun_out/r2_t16.log; grep -v "^ \|Trace\|raise\|check" gpurun_out/r2_ab16.log; cat gpurun_out/r2_explore16.log; grep "^n=" gpurun_out/r2_trace_attn16.log
