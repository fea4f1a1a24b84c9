#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/s4_pytest.txt
( VPFP_PASS2_PREFETCH=2 VPFP_ROWFFT_L2PF=0 timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/s4_pytest_pf2.txt
for pf in 0 1 2; do
  echo "== VPFP_ROWFFT_L2PF=$pf" >> gpurun_out/s4_rowfft.txt
  VPFP_ROWFFT_L2PF=$pf timeout 300 python tools/time_ops.py 16384 16384 "edfdv_exp(table)" 2>&1 | tail -2 >> gpurun_out/s4_rowfft.txt
done
for pf in 0 2; do
  echo "== VPFP_PASS2_PREFETCH=$pf" >> gpurun_out/s4_pass2.txt
  VPFP_PASS2_PREFETCH=$pf timeout 300 python tools/time_ops.py 16384 16384 "vdfdx_exp(table),fp_fast" 2>&1 | tail -4 >> gpurun_out/s4_pass2.txt
done
VPFP_PASS2_PREFETCH=2 VPFP_ROWFFT_L2PF=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s4_bench_pf2.json 2> gpurun_out/s4_bench.err
VPFP_ROWFFT_L2PF=0 VPFP_PASS2_PREFETCH=2 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active --clock-control none \
  -k regex:"rowfft_kernel|pass2_kernel|fp_reg_kernel" -c 3 --csv --log-file gpurun_out/s4_metrics.csv python tools/prof_one.py 16384 16384 all 1 > /dev/null 2>&1
ls -la gpurun_out
