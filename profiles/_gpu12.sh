timeout 900 python -m pytest tests/test_gpu_full.py tests/test_gpu_model.py tests/test_gpu_configs.py tests/test_gpu_encoder.py -m gpu -q -s 2>&1 | grep -v "^\[INFO\]" | grep -E "rel-L2|PSNR|passed|failed|Error|eps|latent|encoder|KL|VQ" > gpurun_out/r2_parity_full.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_t12.log
python profiles/profile_step.py --batch 64 > gpurun_out/r2_plain_step2.log 2>&1 &&
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"implicit_gemm|flash_attention" --launch-count 16 -o gpurun_out/r2_top_kernels -f python profiles/profile_step.py --batch 64 > gpurun_out/r2_ncu_full.log 2>&1
python profiles/one_gn.py > gpurun_out/r2_plain_gn.log 2>&1 &&
timeout 600 ncu --set full --clock-control none -k regex:"gn_" -s 8 -c 4 -o gpurun_out/r2_gn_kernels -f python profiles/one_gn.py > gpurun_out/r2_ncu_gn.log 2>&1
tail -3 gpurun_out/r2_t12.log; tail -2 gpurun_out/r2_ncu_full.log; tail -2 gpurun_out/r2_ncu_gn.log; cat gpurun_out/r2_parity_full.txt | head -60
