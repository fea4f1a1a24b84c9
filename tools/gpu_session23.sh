#!/bin/bash
# A/B: v12 build (base) against the FP kernel with two more barriers removed, with (new) and without (p0) the
# mirror-pair monomial sums; parity subset; then the round's bench line and launch list with the new build
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
OPS="fp_fast+mom,fp_fast_dg"
BASE=$PWD/vlapy_b200/lib/libvpfp_b200_base.so
P0=$PWD/vlapy_b200/lib/libvpfp_b200_p0.so
{
echo "== base (v12)"; VPFP_B200_LIB=$BASE timeout 120 python tools/time_ops.py 16384 16384 "$OPS" 2>&1 | tail -3
echo "== new (pairs)"; timeout 120 python tools/time_ops.py 16384 16384 "$OPS" 2>&1 | tail -3
echo "== p0 (no pairs)"; VPFP_B200_LIB=$P0 timeout 120 python tools/time_ops.py 16384 16384 "$OPS" 2>&1 | tail -3
echo "== base (v12) again"; VPFP_B200_LIB=$BASE timeout 120 python tools/time_ops.py 16384 16384 "$OPS" 2>&1 | tail -3
} > gpurun_out/s23_ab.txt
cat gpurun_out/s23_ab.txt
( timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "fp_sizes or nlepw_c2 or collision or fp_and_moments or stored_quantities" 2>&1 | tail -6 ) > gpurun_out/s23_pytest.txt
cat gpurun_out/s23_pytest.txt
