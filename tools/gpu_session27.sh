#!/bin/bash
# A/B: v14 build (base) against the FP kernel with the one-FMA Z recurrence in its forward sweep; FP parity subset
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
BASE=$PWD/vlapy_b200/lib/libvpfp_b200_base.so
{
echo "== base (v14)"; VPFP_B200_LIB=$BASE timeout 60 python tools/time_ops.py 16384 16384 "fp_fast+mom,fp_fast_dg" 2>&1 | tail -2
echo "== new"; timeout 60 python tools/time_ops.py 16384 16384 "fp_fast+mom,fp_fast_dg" 2>&1 | tail -2
} > gpurun_out/s27_ab.txt
cat gpurun_out/s27_ab.txt
( timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "fp_sizes or collision or fp_and_moments or nlepw_c2" 2>&1 | tail -3 ) > gpurun_out/s27_pytest.txt
cat gpurun_out/s27_pytest.txt
