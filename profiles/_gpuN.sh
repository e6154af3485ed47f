#!/bin/bash
# usage: _gpuN.sh N [extra bench args]: BASELINE configs[2] as written (global batch 64, strong scaling) on N GPUs of one box
N=$1; shift
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 3 --warmup 3 "$@" > gpurun_out/r2_bench_c3_n$N.json 2> gpurun_out/r2_bench_c3_n$N.err
cut -c1-400 gpurun_out/r2_bench_c3_n$N.json; tail -3 gpurun_out/r2_bench_c3_n$N.err
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s 2>&1 | tail -12 > gpurun_out/r2_multi_gpu_invariance.txt; cat gpurun_out/r2_multi_gpu_invariance.txt; fi
