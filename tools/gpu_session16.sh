#!/bin/bash
# N-GPU box: multi-GPU tests and the N-GPU bench line
N=${1:-4}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/s16_pytest_multi_${N}.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/s16_bench_${N}gpu.json 2> gpurun_out/s16_bench_${N}gpu.err
ls -la gpurun_out | tail -5
