#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_t22_$i.log; tail -3 gpurun_out/r2_t22_$i.log; done
grep -B40 "short test summary" gpurun_out/r2_t22_1.log | head -80
