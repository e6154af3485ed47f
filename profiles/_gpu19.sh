#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_t19.log
for sw in "LDM_B200_POLY_EXP=2" "LDM_B200_POLY_EXP=0" "LDM_B200_POLY_EXP=1"; do env $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab19.log 2>&1; done
for sw in "LDM_B200_POLY_EXP=2" "LDM_B200_POLY_EXP=0" "LDM_B200_POLY_EXP=1"; do env AB_B=64 $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab19.log 2>&1; done
timeout 300 python profiles/trace_attn.py > gpurun_out/r2_trace_attn19.log 2>&1
timeout 300 python profiles/explore_batch.py 8 64 > gpurun_out/r2_explore19.log 2>&1
ONLY_AUTO=1 timeout 300 python profiles/trace_epilogue.py > gpurun_out/r2_trace_epi19.log 2>&1
tail -4 gpurun_out/r2_t19.log; grep -v "^ \|Trace\|raise\|check" gpurun_out/r2_ab19.log; cat gpurun_out/r2_explore19.log; grep "^n=" gpurun_out/r2_trace_attn19.log; grep "lean\|FF2\|GEGLU (row" gpurun_out/r2_trace_epi19.log | cut -c1-250
