#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/s10_pytest.txt
for v in 0 1; do
  echo "== VPFP_ROWFFT4=$v" >> gpurun_out/s10_rowfft.txt
  VPFP_ROWFFT4=$v timeout 300 python tools/time_ops.py 16384 16384 "edfdv_exp(table)" 2>&1 | tail -2 >> gpurun_out/s10_rowfft.txt
done
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"rowfft4_kernel" -c 1 \
  -f -o gpurun_out/s10_full python tools/prof_one.py 16384 16384 edfdv 1 > gpurun_out/s10_ncu.log 2>&1
ncu -i gpurun_out/s10_full.ncu-rep --page raw --csv > gpurun_out/s10_full_raw.csv 2>/dev/null
ncu -i gpurun_out/s10_full.ncu-rep --page source --csv --print-source sass > gpurun_out/s10_rowfft4_src.csv 2>/dev/null
rm -f gpurun_out/s10_full.ncu-rep
ls -la gpurun_out
