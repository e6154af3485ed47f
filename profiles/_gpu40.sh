#!/bin/bash
# final sanity of the round's last build (comm.cu rebuilt; rendezvous + checkpoint-reader changes on the host side):
# smoke, the op / tiny-model suite, the full-size parity tests, encoders, properties.  Bounded to ~200 s.
mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke40.log 2>&1
timeout 90 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_t40a.log
timeout 100 python -m pytest tests/test_gpu_full.py tests/test_gpu_encoder.py tests/test_gpu_properties.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_t40b.log
tail -2 gpurun_out/r2_smoke40.log; cat gpurun_out/r2_t40a.log gpurun_out/r2_t40b.log
