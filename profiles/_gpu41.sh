#!/bin/bash
# the shim's new shape validation against the real call sites: tiny-model suite, encoders, K5 / K6 op tests (bounded to ~75 s)
mkdir -p gpurun_out
timeout 75 python -m pytest tests/test_gpu_model.py tests/test_gpu_encoder.py tests/test_gpu_properties.py tests/test_gpu_ops.py -m gpu -x -q -k "not test_full_size_encoder_256" 2>&1 | tail -6 > gpurun_out/r2_t41.log
cat gpurun_out/r2_t41.log
