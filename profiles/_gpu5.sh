timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_t5.log
timeout 300 python profiles/trace_epilogue.py > gpurun_out/r2_trace_epi3.log 2>&1
for sw in "X=1" "LDM_B200_W16=0" "LDM_B200_W16=0 LDM_B200_EW4=0" "LDM_B200_FRAG_GEGLU=1"; do env $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab5.log 2>&1; done
for sw in "X=1" "LDM_B200_W16=0"; do env AB_B=64 $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab5.log 2>&1; done
timeout 300 python profiles/gemm_shapes.py > gpurun_out/r2_gemm_shapes_b8_v3.log 2>&1
tail -4 gpurun_out/r2_t5.log; grep -v "^ \|Trace\|raise\|check" gpurun_out/r2_ab5.log
