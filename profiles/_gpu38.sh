#!/bin/bash
# re-entry sanity of the rebuilt library (fresh container): smoke + the op / tiny-model parity tests, bounded to ~5 min
mkdir -p gpurun_out
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke38.log 2>&1
timeout 200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_t38.log
tail -3 gpurun_out/r2_smoke38.log; tail -3 gpurun_out/r2_t38.log
