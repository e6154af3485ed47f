timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_t8.log
for sw in "LDM_B200_LEAN_EW=8" "LDM_B200_LEAN_EW=12" "LDM_B200_LEAN_EW=16" "LDM_B200_POLY_EXP=0" "LDM_B200_W16_MINSTAGES=99"; do env $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab8.log 2>&1; done
for sw in "LDM_B200_LEAN_EW=8" "LDM_B200_LEAN_EW=12" "LDM_B200_LEAN_EW=16" "LDM_B200_POLY_EXP=0" "LDM_B200_W16_MINSTAGES=99"; do env AB_B=64 $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab8.log 2>&1; done
for sw in "LDM_B200_LEAN_EW=12" "LDM_B200_LEAN_EW=16"; do echo "== $sw" >> gpurun_out/r2_trace_ew.log; env ONLY_AUTO=1 $sw timeout 300 python profiles/trace_epilogue.py >> gpurun_out/r2_trace_ew.log 2>&1; done
timeout 300 python profiles/gemm_shapes.py > gpurun_out/r2_gemm_shapes_b8_v6.log 2>&1
tail -4 gpurun_out/r2_t8.log; grep -v "^ \|Trace\|raise\|check" gpurun_out/r2_ab8.log
