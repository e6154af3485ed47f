#!/bin/bash
mkdir -p gpurun_out
for sw in "LDM_B200_T_LEAN4_MAXKB=1000" "LDM_B200_T_LEAN4_MAXKB=20" "LDM_B200_T_LEAN4_MAXKB=10" "LDM_B200_LEAN_EW4=0" "LDM_B200_T_LEAN4_MAXKB=45"; do env AB_B=64 $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab29.log 2>&1; done
for sw in "LDM_B200_T_LEAN4_MAXKB=1000" "LDM_B200_T_LEAN4_MAXKB=20" "LDM_B200_LEAN_EW4=0"; do env $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab29.log 2>&1; done
SHAPES_B=64 LDM_B200_LEAN_EW4=1 timeout 300 python profiles/gemm_shapes.py > gpurun_out/r2_gemm_shapes29_on.log 2>&1
SHAPES_B=64 LDM_B200_LEAN_EW4=0 timeout 300 python profiles/gemm_shapes.py > gpurun_out/r2_gemm_shapes29_off.log 2>&1
grep -v "^ \|Trace\|raise\|check" gpurun_out/r2_ab29.log; head -30 gpurun_out/r2_gemm_shapes29_on.log | cut -c1-150; echo; head -30 gpurun_out/r2_gemm_shapes29_off.log | cut -c1-150
