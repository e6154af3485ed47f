#!/bin/bash
# validation mode tests
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_validate.py -m gpu -x -q -s > gpurun_out/r2_t15.log 2>&1
tail -25 gpurun_out/r2_t15.log
