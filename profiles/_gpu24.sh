#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t24.log
timeout 900 python bench.py > gpurun_out/r2_bench_c3_n1.json 2> gpurun_out/r2_bench_c3_n1.err
timeout 900 python bench.py --config c2 --steps 2 --warmup 3 > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err
timeout 900 python bench.py --config c4 --steps 2 --warmup 3 > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err
timeout 900 python bench.py --config c5 --steps 3 --warmup 3 > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_bench_c5.err
timeout 300 python profiles/ab_step.py > gpurun_out/r2_ab24.log 2>&1
AB_B=64 timeout 300 python profiles/ab_step.py >> gpurun_out/r2_ab24.log 2>&1
tail -3 gpurun_out/r2_t24.log; for c in c3_n1 c2 c4 c5; do cut -c1-160 gpurun_out/r2_bench_$c.json; tail -2 gpurun_out/r2_bench_$c.err; done; grep -v "^ \|Trace\|raise\|check" gpurun_out/r2_ab24.log
