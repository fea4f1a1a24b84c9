#!/bin/bash
# compute-sanitizer racecheck + memcheck on the kernels that rely on own-slot cp.async prefetch and in-place exchanges
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
SEL="test_fp_sizes_vs_oracle and (4096 or 16384) and True or test_single_pass_row_kernel and 257 or test_vdfdx_fused_density and 4096 or test_fast_and_generic_kernels_agree_with_oracle and 4096-1"
( timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" 2>&1 | tail -40 ) > gpurun_out/s17_racecheck.txt
( timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" 2>&1 | tail -25 ) > gpurun_out/s17_memcheck.txt
ls -la gpurun_out | tail -3
