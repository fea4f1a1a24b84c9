#!/bin/bash
# 2-GPU box: full GPU test suite (incl. multi-GPU), 1- and 2-GPU bench lines
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/s14_pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s14_bench_1gpu.json 2> gpurun_out/s14_bench_1gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/s14_bench_2gpu.json 2> gpurun_out/s14_bench_2gpu.err
timeout 300 python __graft_entry__.py smoke > gpurun_out/s14_smoke.txt 2>&1
ls -la gpurun_out
