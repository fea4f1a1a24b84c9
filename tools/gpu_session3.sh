#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/s3_pytest.txt
for cfg in "VPFP_FP_REG_M=32" "VPFP_FP_REG_M=32 VPFP_FP_BURST=1" "VPFP_FP_REG_M=64"; do
  echo "== $cfg" >> gpurun_out/s3_fp_ab.txt
  env $cfg timeout 300 python tools/time_ops.py 16384 16384 fp_fast 2>&1 | tail -3 >> gpurun_out/s3_fp_ab.txt
done
for nv in 256 512 1024 2048 4096; do
  echo "== 16384 x $nv" >> gpurun_out/s3_vdfdx_l2.txt
  timeout 300 python tools/time_ops.py 16384 $nv "vdfdx_exp(table),copy,edfdv_exp(table)" 2>&1 | tail -3 >> gpurun_out/s3_vdfdx_l2.txt
done
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"fp_reg_kernel" -c 1 \
  -f -o gpurun_out/s3_full python tools/prof_one.py 16384 16384 fp 1 > gpurun_out/s3_ncu.log 2>&1
ncu -i gpurun_out/s3_full.ncu-rep --page raw --csv > gpurun_out/s3_full_raw.csv 2>/dev/null
ncu -i gpurun_out/s3_full.ncu-rep --page source --csv --print-source sass > gpurun_out/s3_fp_src.csv 2>/dev/null
rm -f gpurun_out/s3_full.ncu-rep
ls -la gpurun_out
