#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( VPFP_PASS2_PREFETCH=2 VPFP_ROWFFT_L2PF=0 timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/s6_pytest.txt
( VPFP_PASS2_PREFETCH=0 VPFP_ROWFFT_L2PF=0 timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/s6_pytest_pf0.txt
for pf in 0 1; do
  echo "== VPFP_ROWFFT_L2PF=$pf" >> gpurun_out/s6_rowfft.txt
  VPFP_ROWFFT_L2PF=$pf timeout 300 python tools/time_ops.py 16384 16384 "edfdv_exp(table)" 2>&1 | tail -2 >> gpurun_out/s6_rowfft.txt
done
for pf in 0 2; do
  echo "== VPFP_PASS2_PREFETCH=$pf" >> gpurun_out/s6_pass2.txt
  VPFP_PASS2_PREFETCH=$pf timeout 300 python tools/time_ops.py 16384 16384 "vdfdx_exp(table),xmodes" 2>&1 | tail -3 >> gpurun_out/s6_pass2.txt
done
VPFP_PASS2_PREFETCH=2 VPFP_ROWFFT_L2PF=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err
ls -la gpurun_out
