#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_t31.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke31.log 2>&1
timeout 900 python bench.py > gpurun_out/r2_bench_c3_n1.json 2> gpurun_out/r2_bench_c3_n1.err
tail -3 gpurun_out/r2_t31.log; tail -3 gpurun_out/r2_smoke31.log; cut -c1-200 gpurun_out/r2_bench_c3_n1.json; tail -2 gpurun_out/r2_bench_c3_n1.err
