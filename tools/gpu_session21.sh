#!/bin/bash
# round-1 v11 measurements: bench line (with the CPU baseline), ncu launch list of the bench command, ncu --set full of
# the operators (DRAM traffic per kernel for roofline.traffic), smoke
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/s21_bench_1gpu.json 2> gpurun_out/s21_bench_1gpu.err
cut -c1-300 gpurun_out/s21_bench_1gpu.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s21_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/s21_launches_bench.log 2>&1
python tools/ncu_summary.py gpurun_out/s21_launches.csv > gpurun_out/s21_launches.txt 2>&1; head -12 gpurun_out/s21_launches.txt
timeout 300 ncu --set full --clock-control none \
  -k regex:"rowfft_kernel|fp_reg_kernel|pass13_kernel|pass2_kernel|Xmodes" -c 6 \
  -f -o gpurun_out/s21_full python tools/prof_one.py 16384 16384 all 1 > gpurun_out/s21_ncu.log 2>&1
ncu -i gpurun_out/s21_full.ncu-rep --page raw --csv > gpurun_out/s21_full_raw.csv 2>/dev/null
python tools/ncu_traffic.py gpurun_out/s21_full_raw.csv gpurun_out/s21_traffic.json gpurun_out/s21_full.txt "round 1 v11 kernels (tools/prof_one.py 16384 16384 all 1)" | grep -E "^==|gpu__time|dram__bytes|fp64"
rm -f gpurun_out/s21_full.ncu-rep
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
