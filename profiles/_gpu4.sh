python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_t4.log
python profiles/trace_epilogue.py > gpurun_out/r2_trace_epi2.log 2>&1
python profiles/ab_step.py >> gpurun_out/r2_ab4.log 2>&1
LDM_B200_FRAG_GEGLU=1 python profiles/ab_step.py >> gpurun_out/r2_ab4.log 2>&1
LDM_B200_FRAG16=1 python profiles/ab_step.py >> gpurun_out/r2_ab4.log 2>&1
LDM_B200_EW4=0 python profiles/ab_step.py >> gpurun_out/r2_ab4.log 2>&1
AB_B=64 python profiles/ab_step.py >> gpurun_out/r2_ab4.log 2>&1
AB_B=64 LDM_B200_FRAG_GEGLU=1 python profiles/ab_step.py >> gpurun_out/r2_ab4.log 2>&1
AB_B=64 LDM_B200_EW4=0 python profiles/ab_step.py >> gpurun_out/r2_ab4.log 2>&1
python profiles/gemm_shapes.py > gpurun_out/r2_gemm_shapes_b8_v2.log 2>&1
tail -4 gpurun_out/r2_t4.log; cat gpurun_out/r2_ab4.log
