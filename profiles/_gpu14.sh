timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t14.log
for sw in "X=1" "LDM_B200_GN_VEC8=0"; do env $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab14.log 2>&1; done
for sw in "X=1" "LDM_B200_GN_VEC8=0"; do env AB_B=64 $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab14.log 2>&1; done
timeout 300 python profiles/explore_batch.py 8 16 32 64 > gpurun_out/r2_explore14.log 2>&1
tail -3 gpurun_out/r2_t14.log; grep -v "^ \|Trace\|raise\|check" gpurun_out/r2_ab14.log; cat gpurun_out/r2_explore14.log
