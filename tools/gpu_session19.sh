#!/bin/bash
# instruction-level ncu capture (source counters, per-instruction stall samples) of the three compute-bound
# kernels of the step: e df/dv row kernel, Fokker-Planck kernel, v df/dx pass 2
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 400 ncu --set full --clock-control none --import-source on \
  -k regex:"rowfft_kernel|fp_reg_kernel|pass2_kernel" -c 3 \
  -f -o gpurun_out/s19_src python tools/prof_one.py 16384 16384 all 1 > gpurun_out/s19_ncu.log 2>&1
tail -3 gpurun_out/s19_ncu.log
ls -la gpurun_out | tail -4
