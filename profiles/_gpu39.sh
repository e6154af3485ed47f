#!/bin/bash
# re-entry sanity, second half: the full-size parity / config / encoder / property / validation-mode tests (bounded)
mkdir -p gpurun_out
timeout 340 python -m pytest tests/test_gpu_full.py tests/test_gpu_configs.py tests/test_gpu_encoder.py tests/test_gpu_properties.py tests/test_gpu_validate.py -m gpu -x -q --durations=8 2>&1 | tail -16 > gpurun_out/r2_t39.log
cat gpurun_out/r2_t39.log
