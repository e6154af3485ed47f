echo "== lean, fine stamps" > gpurun_out/r2_trace_lean_fine.log; LDM_B200_TRACE_FINE=1 ONLY_AUTO=1 timeout 300 python profiles/trace_epilogue.py >> gpurun_out/r2_trace_lean_fine.log 2>&1
for ew in 8 12 16; do echo "== accumulator drain only (TMEM -> registers), $ew epilogue warps" >> gpurun_out/r2_trace_tmem_only.log; LDM_B200_LEAN_EW=$ew LDM_B200_TRACE_TMEM_ONLY=1 ONLY_AUTO=1 timeout 300 python profiles/trace_epilogue.py >> gpurun_out/r2_trace_tmem_only.log 2>&1; done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_c3_n1.json 2> gpurun_out/r2_bench_c3_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
cat gpurun_out/r2_smoke.log; cut -c1-300 gpurun_out/r2_bench_c3_n1.json; cut -c1-300 gpurun_out/r2_bench_ref.json; grep "C x C, 16-bit out (row-owner lean)" gpurun_out/r2_trace_tmem_only.log | cut -c1-250
