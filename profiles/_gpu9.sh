timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_t9.log
for sw in "X=1" "LDM_B200_LEAN=0"; do env $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab9.log 2>&1; done
for sw in "X=1"; do env AB_B=64 $sw timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab9.log 2>&1; done
timeout 300 python profiles/gemm_shapes.py > gpurun_out/r2_gemm_shapes_b8_v7.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench_c3_v2.json 2> gpurun_out/r2_bench_c3_v2.err
timeout 600 python bench.py --config c5 --steps 2 --warmup 3 > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_bench_c5.err
tail -4 gpurun_out/r2_t9.log; grep -v "^ \|Trace\|raise\|check" gpurun_out/r2_ab9.log; cut -c1-600 gpurun_out/r2_bench_c3_v2.json; tail -3 gpurun_out/r2_bench_c5.err
