#!/bin/bash
# A/B: v11 build (lib/libvpfp_b200_base.so) against the new build (rowfft without the barriers after its pointwise and
# last phases, FP kernel without the end-of-row barrier, density reduction with eight loads in flight) + parity subset
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
OPS="edfdv_exp(table),vdfdx_exp(table),fp_fast+mom"
BASE=$PWD/vlapy_b200/lib/libvpfp_b200_base.so
{
echo "== base (v11)"; VPFP_B200_LIB=$BASE timeout 120 python tools/time_ops.py 16384 16384 "$OPS" 2>&1 | tail -4
echo "== new"; timeout 120 python tools/time_ops.py 16384 16384 "$OPS" 2>&1 | tail -4
echo "== base (v11) again"; VPFP_B200_LIB=$BASE timeout 120 python tools/time_ops.py 16384 16384 "$OPS" 2>&1 | tail -4
echo "== new again"; timeout 120 python tools/time_ops.py 16384 16384 "$OPS" 2>&1 | tail -4
} > gpurun_out/s22_ab.txt
cat gpurun_out/s22_ab.txt
( timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_pass or fused_density or fp_sizes or nlepw_c2 or edfdv_sizes" 2>&1 | tail -6 ) > gpurun_out/s22_pytest.txt
cat gpurun_out/s22_pytest.txt
