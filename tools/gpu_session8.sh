#!/bin/bash
# 2-GPU session: multi-GPU tests (peer scatter through the new kernels) and the 2-GPU bench line
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi -L > gpurun_out/s8_gpus.txt
( timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/s8_pytest_multi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/s8_bench_2gpu.json 2> gpurun_out/s8_bench_2gpu.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s8_bench_1gpu.json 2> gpurun_out/s8_bench_1gpu.err
ls -la gpurun_out
