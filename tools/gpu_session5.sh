#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( VPFP_PASS2_PREFETCH=2 VPFP_ROWFFT_L2PF=0 timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/s5_pytest.txt
for pf in 0 1; do
  echo "== VPFP_ROWFFT_L2PF=$pf" >> gpurun_out/s5_rowfft.txt
  VPFP_ROWFFT_L2PF=$pf timeout 300 python tools/time_ops.py 16384 16384 "edfdv_exp(table)" 2>&1 | tail -2 >> gpurun_out/s5_rowfft.txt
done
VPFP_ROWFFT_L2PF=0 timeout 300 python tools/time_ops.py 8192 8192 "edfdv_exp(table),vdfdx_exp(table),fp_fast,copy" > gpurun_out/s5_8192.txt 2>&1
VPFP_ROWFFT_L2PF=0 timeout 300 python tools/time_ops.py 4096 4096 "edfdv_exp(table),vdfdx_exp(table),fp_fast,copy" > gpurun_out/s5_4096.txt 2>&1
VPFP_ROWFFT_L2PF=0 timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"rowfft_kernel" -c 1 \
  -f -o gpurun_out/s5_full python tools/prof_one.py 16384 16384 edfdv 1 > gpurun_out/s5_ncu.log 2>&1
ncu -i gpurun_out/s5_full.ncu-rep --page raw --csv > gpurun_out/s5_full_raw.csv 2>/dev/null
ncu -i gpurun_out/s5_full.ncu-rep --page source --csv --print-source sass > gpurun_out/s5_rowfft_src.csv 2>/dev/null
rm -f gpurun_out/s5_full.ncu-rep
ls -la gpurun_out
