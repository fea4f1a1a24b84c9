#!/bin/bash
# GPU session: parity with the register-resident FP kernel, FP A/B, v df/dx L2-slab sweep, ncu --set full
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/s1_gpu.txt
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/s1_pytest.txt
for cfg in VPFP_NO_FP_REG=1 VPFP_FP_REG_M=32 VPFP_FP_REG_M=64; do
  echo "== $cfg" >> gpurun_out/s1_fp_ab.txt
  env $cfg timeout 300 python tools/time_ops.py 16384 16384 fp_fast 2>&1 | tail -4 >> gpurun_out/s1_fp_ab.txt
done
for mb in 0 16 32 64; do for ns in 1 3; do
  echo "== VPFP_SLAB_MB=$mb VPFP_SLAB_STREAMS=$ns" >> gpurun_out/s1_slab.txt
  VPFP_SLAB_MB=$mb VPFP_SLAB_STREAMS=$ns timeout 300 python tools/time_ops.py 16384 16384 "vdfdx_exp(table)" 2>&1 | tail -2 >> gpurun_out/s1_slab.txt
done; done
timeout 300 python tools/time_ops.py 16384 16384 > gpurun_out/s1_time_ops.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"rowfft_kernel|fp_reg_kernel|pass13_kernel|pass2_kernel|XmodesProg" -c 7 \
  -f -o gpurun_out/s1_full python tools/prof_one.py 16384 16384 all 1 > gpurun_out/s1_ncu.log 2>&1
ncu -i gpurun_out/s1_full.ncu-rep --page raw --csv > gpurun_out/s1_full_raw.csv 2>/dev/null
ls -la gpurun_out
