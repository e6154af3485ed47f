python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_t3.log
python profiles/trace_epilogue.py > gpurun_out/r2_trace_epi.log 2>&1
python profiles/gemm_shapes.py > gpurun_out/r2_gemm_shapes_b8.log 2>&1
python profiles/ab_step.py >> gpurun_out/r2_ab3.log 2>&1
AB_B=64 python profiles/ab_step.py >> gpurun_out/r2_ab3.log 2>&1
tail -4 gpurun_out/r2_t3.log; cat gpurun_out/r2_ab3.log
