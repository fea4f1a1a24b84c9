#!/bin/bash
# A/B of two builds of the library in one session: lib/libvpfp_b200_base.so (previous commit) against the new build
# (row tables prepared a phase early in rowfft.cuh, pass 2 tile loop without spilled loop bounds, FP moment sums
# interleaved with the transposed stores), plus pass-2 tiles per CTA; then the parity tests that cover the three kernels
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
OPS="edfdv_exp(table),vdfdx_exp(table),fp_fast+mom"
BASE=$PWD/vlapy_b200/lib/libvpfp_b200_base.so
{
echo "== base"; VPFP_B200_LIB=$BASE timeout 120 python tools/time_ops.py 16384 16384 "$OPS" 2>&1 | tail -4
echo "== new"; timeout 120 python tools/time_ops.py 16384 16384 "$OPS" 2>&1 | tail -4
echo "== new, VPFP_PASS2_CHUNK=16"; VPFP_PASS2_CHUNK=16 timeout 120 python tools/time_ops.py 16384 16384 "vdfdx_exp(table)" 2>&1 | tail -2
echo "== new, VPFP_PASS2_CHUNK=4"; VPFP_PASS2_CHUNK=4 timeout 120 python tools/time_ops.py 16384 16384 "vdfdx_exp(table)" 2>&1 | tail -2
echo "== base again"; VPFP_B200_LIB=$BASE timeout 120 python tools/time_ops.py 16384 16384 "$OPS" 2>&1 | tail -4
} > gpurun_out/s20_ab.txt
cat gpurun_out/s20_ab.txt
( timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_pass or fast_and_generic or fused_density or fp_sizes or ensemble_with_per or nlepw_c2" 2>&1 | tail -6 ) > gpurun_out/s20_pytest.txt
cat gpurun_out/s20_pytest.txt
