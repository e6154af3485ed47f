timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_t7.log
for sw in "X=1" "LDM_B200_W16=0" "LDM_B200_LEAN=0"; do echo "== $sw" >> gpurun_out/r2_trace_lean.log; env ONLY_AUTO=1 $sw timeout 300 python profiles/trace_epilogue.py >> gpurun_out/r2_trace_lean.log 2>&1; done
for sw in "X=1" "LDM_B200_W16=0" "LDM_B200_LEAN=0" "LDM_B200_STREAM=fp32"; do env $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab7.log 2>&1; done
for sw in "X=1" "LDM_B200_W16=0" "LDM_B200_LEAN=0" "LDM_B200_STREAM=fp32"; do env AB_B=64 $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab7.log 2>&1; done
timeout 300 python profiles/gemm_shapes.py > gpurun_out/r2_gemm_shapes_b8_v5.log 2>&1
tail -4 gpurun_out/r2_t7.log; grep -v "^ \|Trace\|raise\|check" gpurun_out/r2_ab7.log
