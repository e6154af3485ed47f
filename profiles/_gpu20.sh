#!/bin/bash
# final captures of the round: bench line, parity numbers, ncu launch list + DRAM bytes, ncu --set full of the top kernels and K2
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2_bench_c3_n1.json 2> gpurun_out/r2_bench_c3_n1.err
timeout 900 python -m pytest tests/test_gpu_full.py tests/test_gpu_model.py tests/test_gpu_configs.py tests/test_gpu_encoder.py tests/test_gpu_validate.py -m gpu -q -s 2>&1 | grep -v "^\[INFO\]" | grep -E "rel-L2|PSNR|passed|failed|Error|eps|latent|encoder|KL|VQ" > gpurun_out/r2_parity_full.txt
python profiles/profile_step.py --batch 64 > gpurun_out/r2_plain_step.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_dram_unet_step_b64.csv python profiles/profile_step.py --batch 64 > gpurun_out/r2_ncu_step.log 2>&1
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"implicit_gemm|flash_attention" --launch-count 16 -o gpurun_out/r2_top_kernels -f python profiles/profile_step.py --batch 64 > gpurun_out/r2_ncu_full.log 2>&1
python profiles/one_gn.py > gpurun_out/r2_plain_gn.log 2>&1 &&
timeout 600 ncu --set full --clock-control none -k regex:"gn_" -s 8 -c 4 -o gpurun_out/r2_gn_kernels -f python profiles/one_gn.py > gpurun_out/r2_ncu_gn.log 2>&1
cut -c1-600 gpurun_out/r2_bench_c3_n1.json; tail -3 gpurun_out/r2_bench_c3_n1.err; tail -2 gpurun_out/r2_ncu_full.log; tail -2 gpurun_out/r2_ncu_gn.log; cat gpurun_out/r2_plain_gn.log | tail -2; tail -5 gpurun_out/r2_parity_full.txt
