python -m pytest tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2_t2_ops.log
python -m pytest tests/test_gpu_model.py tests/test_gpu_full.py tests/test_gpu_configs.py tests/test_gpu_properties.py -m gpu -x -q -s 2>&1 | grep -v "^\[INFO\]" | tail -80 > gpurun_out/r2_t2_model.log
python profiles/explore_batch.py 8 64 > gpurun_out/r2_explore2.log 2>&1
for sw in "X=1" "LDM_B200_FRAG16=0" "LDM_B200_UPCONV=materialize" "LDM_B200_DOWNCONV=im2col"; do env $sw python profiles/ab_step.py >> gpurun_out/r2_ab2.log 2>&1; done
for sw in "X=1" "LDM_B200_FRAG16=0"; do env AB_B=64 $sw python profiles/ab_step.py >> gpurun_out/r2_ab2.log 2>&1; done
tail -3 gpurun_out/r2_t2_ops.log; tail -3 gpurun_out/r2_t2_model.log; cat gpurun_out/r2_explore2.log gpurun_out/r2_ab2.log
