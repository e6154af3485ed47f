for sw in "X=1" "LDM_B200_T_EPI_COST=15" "LDM_B200_T_EPI_COST=25" "LDM_B200_T_TILE_OVH=800" "LDM_B200_T_EPI_COST=20 LDM_B200_T_TILE_OVH=800" "LDM_B200_PDL_MASK=3" "LDM_B200_PDL_MASK=7"; do env $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab11.log 2>&1; done
for sw in "X=1" "LDM_B200_T_EPI_COST=20" "LDM_B200_PDL_MASK=3"; do env AB_B=64 $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab11.log 2>&1; done
timeout 900 python bench.py --config c2 --steps 2 --warmup 3 > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err
timeout 900 python bench.py --config c4 --steps 2 --warmup 3 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err
python profiles/profile_step.py --batch 64 > gpurun_out/r2_plain_step.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_dram_unet_step_b64.csv python profiles/profile_step.py --batch 64 > gpurun_out/r2_ncu_step.log 2>&1
grep -v "^ \|Trace\|raise\|check" gpurun_out/r2_ab11.log; cut -c1-200 gpurun_out/r2_bench_c2.json; cut -c1-200 gpurun_out/r2_bench_c4.json; tail -3 gpurun_out/r2_ncu_step.log
