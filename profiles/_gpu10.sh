# 2 GPUs: multi-GPU invariance through the library's NCCL communicator, then the strong-scaling bench at N = 2
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s 2>&1 | tail -15 > gpurun_out/r2_t10_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 2 --verify --no-rooflines > gpurun_out/r2_bench_c3_n2.json 2> gpurun_out/r2_bench_c3_n2.err
tail -5 gpurun_out/r2_t10_multi.log; cut -c1-300 gpurun_out/r2_bench_c3_n2.json; tail -5 gpurun_out/r2_bench_c3_n2.err
