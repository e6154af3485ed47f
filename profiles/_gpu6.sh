timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_t6.log
LDM_B200_TRACE_FINE=1 timeout 300 python profiles/trace_epilogue.py > gpurun_out/r2_trace_fine.log 2>&1
for sw in "X=1" "LDM_B200_STREAM=fp32" "LDM_B200_W16=0"; do env $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab6.log 2>&1; done
for sw in "X=1" "LDM_B200_STREAM=fp32" "LDM_B200_W16=0"; do env AB_B=64 $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab6.log 2>&1; done
timeout 300 python profiles/explore_batch.py 8 64 > gpurun_out/r2_explore6.log 2>&1
timeout 300 python profiles/gemm_shapes.py > gpurun_out/r2_gemm_shapes_b8_v4.log 2>&1
tail -4 gpurun_out/r2_t6.log; grep -v "^ \|Trace\|raise\|check" gpurun_out/r2_ab6.log; cat gpurun_out/r2_explore6.log
