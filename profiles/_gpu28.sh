#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2_t28.log
for sw in "LDM_B200_LEAN_EW4=1" "LDM_B200_LEAN_EW4=0"; do env $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab28.log 2>&1; done
for sw in "LDM_B200_LEAN_EW4=1" "LDM_B200_LEAN_EW4=0"; do env AB_B=64 $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab28.log 2>&1; done
for sw in "LDM_B200_LEAN_EW4=1" "LDM_B200_LEAN_EW4=0"; do env AB_B=16 $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab28.log 2>&1; done
timeout 300 python profiles/gemm_shapes.py > gpurun_out/r2_gemm_shapes28.log 2>&1
tail -4 gpurun_out/r2_t28.log; grep -v "^ \|Trace\|raise\|check" gpurun_out/r2_ab28.log; head -12 gpurun_out/r2_gemm_shapes28.log | cut -c1-150
