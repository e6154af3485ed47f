#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke36.log 2>&1
timeout 900 python bench.py > gpurun_out/r2_bench_c3_n1.json 2> gpurun_out/r2_bench_c3_n1.err
timeout 600 python -m pytest tests/test_gpu_validate.py -m gpu -q -s 2>&1 | grep -E "validation|passed|failed" > gpurun_out/r2_t15.log
tail -3 gpurun_out/r2_smoke36.log; cut -c1-200 gpurun_out/r2_bench_c3_n1.json; tail -2 gpurun_out/r2_bench_c3_n1.err; tail -3 gpurun_out/r2_t15.log
