#!/bin/bash
# GPU session 2: FP kernel variants (coalesced stores, prefetch pattern, M), rowfft L2 prefetch, fused density
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/s2_pytest.txt
for cfg in "VPFP_FP_REG_M=32" "VPFP_FP_REG_M=32 VPFP_FP_BURST=1" "VPFP_FP_REG_M=64" "VPFP_FP_REG_M=64 VPFP_FP_BURST=1"; do
  echo "== $cfg" >> gpurun_out/s2_fp_ab.txt
  env $cfg timeout 300 python tools/time_ops.py 16384 16384 fp_fast 2>&1 | tail -3 >> gpurun_out/s2_fp_ab.txt
done
for pf in 0 1 2 4; do
  echo "== VPFP_ROWFFT_L2PF=$pf" >> gpurun_out/s2_rowfft.txt
  VPFP_ROWFFT_L2PF=$pf timeout 300 python tools/time_ops.py 16384 16384 "edfdv_exp(table)" 2>&1 | tail -2 >> gpurun_out/s2_rowfft.txt
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"fp_reg_kernel|prog_kernel" -c 3 \
  -f -o gpurun_out/s2_full python tools/prof_one.py 16384 16384 fpx 1 > gpurun_out/s2_ncu.log 2>&1
ncu -i gpurun_out/s2_full.ncu-rep --page raw --csv > gpurun_out/s2_full_raw.csv 2>/dev/null
VPFP_ROWFFT_L2PF=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:"rowfft_kernel" -c 2 --csv --log-file gpurun_out/s2_rowfft_nopf.csv python tools/prof_one.py 16384 16384 edfdv 1 > /dev/null 2>&1
ls -la gpurun_out
