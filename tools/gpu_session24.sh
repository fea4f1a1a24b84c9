#!/bin/bash
# round-1 v13: FP kernel A/B against the v12 build, L2-prefetch knob of the row kernel, bench line + launch list
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
BASE=$PWD/vlapy_b200/lib/libvpfp_b200_base.so
{
echo "== base (v12)"; VPFP_B200_LIB=$BASE timeout 120 python tools/time_ops.py 16384 16384 "fp_fast+mom,fp_fast_dg" 2>&1 | tail -3
echo "== new"; timeout 120 python tools/time_ops.py 16384 16384 "fp_fast+mom,fp_fast_dg,edfdv_exp(table)" 2>&1 | tail -4
echo "== new, VPFP_ROWFFT_L2PF=1"; VPFP_ROWFFT_L2PF=1 timeout 120 python tools/time_ops.py 16384 16384 "edfdv_exp(table)" 2>&1 | tail -2
} > gpurun_out/s24_ab.txt
cat gpurun_out/s24_ab.txt
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/s24_bench_1gpu.json 2> gpurun_out/s24_bench_1gpu.err
cut -c1-330 gpurun_out/s24_bench_1gpu.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s24_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/s24_launches_bench.log 2>&1
python tools/ncu_summary.py gpurun_out/s24_launches.csv > gpurun_out/s24_launches.txt 2>&1; grep -v "at::" gpurun_out/s24_launches.txt | head -14
