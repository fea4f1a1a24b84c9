#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/s13_pytest.txt
timeout 300 python tools/time_ops.py 16384 16384 "edfdv_exp,vdfdx_exp,fp_fast,xmodes" > gpurun_out/s13_ops.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s13_bench.json 2> gpurun_out/s13_bench.err
ls -la gpurun_out
