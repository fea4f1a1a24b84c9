#!/bin/bash
# First GPU call of the next round.  Prepare on the CPU side first:
#     python tools/make_ab.py HEAD -DFPREG_ZFMA_SPIKE=1        # base = HEAD, candidate = spike sweeps with the one-FMA recurrence
# 1. A/B of the prepared candidates (FP spike sweeps; x-mode CTA width / row chunks);
# 2. instruction-level capture of the three compute-bound kernels as they are now (source page -> tools/ncu_phases.py,
#    tools/ncu_phase_ops.py here); about 3 GPU-minutes in all.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
BASE=$PWD/vlapy_b200/lib/libvpfp_b200_base.so
CAND=$PWD/vlapy_b200/lib/libvpfp_b200_cand.so
{
echo "== base"; VPFP_B200_LIB=$BASE timeout 60 python tools/time_ops.py 16384 16384 "fp_fast+mom,fp_fast_dg,xmodes" 2>&1 | tail -3
echo "== candidate (FPREG_ZFMA_SPIKE=1)"; VPFP_B200_LIB=$CAND timeout 60 python tools/time_ops.py 16384 16384 "fp_fast+mom,fp_fast_dg" 2>&1 | tail -2
for t in 256 512; do echo "== VPFP_XMODES_THREADS=$t"; VPFP_XMODES_THREADS=$t timeout 60 python tools/time_ops.py 16384 16384 "xmodes" 2>&1 | tail -1; done
for c in 16 64 128; do echo "== VPFP_XMODES_XCH=$c"; VPFP_XMODES_XCH=$c timeout 60 python tools/time_ops.py 16384 16384 "xmodes" 2>&1 | tail -1; done
} > gpurun_out/next_ab.txt
cat gpurun_out/next_ab.txt
( VPFP_B200_LIB=$CAND timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fp_sizes or collision or nlepw_c2" 2>&1 | tail -3 ) > gpurun_out/next_pytest_cand.txt
cat gpurun_out/next_pytest_cand.txt
timeout 400 ncu --set full --clock-control none --import-source on \
  -k regex:"rowfft_kernel|fp_reg_kernel|pass2_kernel" -c 3 \
  -f -o gpurun_out/next_src python tools/prof_one.py 16384 16384 all 1 > gpurun_out/next_ncu.log 2>&1
tail -2 gpurun_out/next_ncu.log
