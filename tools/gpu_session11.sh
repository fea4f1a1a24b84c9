#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/s11_pytest.txt
( VPFP_ROWFFT4=1 VPFP_PASS2_PREFETCH=0 timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/s11_pytest_alt.txt
for v in 0 1; do
  echo "== VPFP_ROWFFT4=$v" >> gpurun_out/s11_ops.txt
  VPFP_ROWFFT4=$v timeout 300 python tools/time_ops.py 16384 16384 "edfdv_exp(table),vdfdx_exp(table)" 2>&1 | tail -3 >> gpurun_out/s11_ops.txt
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s11_bench.json 2> gpurun_out/s11_bench.err
ls -la gpurun_out
