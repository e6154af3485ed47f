#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t37.log
timeout 300 python profiles/ab_step.py > gpurun_out/r2_ab37.log 2>&1
AB_B=64 timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab37.log 2>&1
SHAPES_B=64 timeout 300 python profiles/gemm_shapes.py > gpurun_out/r2_gemm_shapes37.log 2>&1
tail -3 gpurun_out/r2_t37.log; grep -v "^ \|Trace\|raise\|check" gpurun_out/r2_ab37.log; head -14 gpurun_out/r2_gemm_shapes37.log | cut -c1-150
