#!/bin/bash
# final measurement session: tests, ncu --set full of the C5 operators -> traffic table, bench line, ncu launch list
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/s7_pytest.txt
timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:"rowfft_kernel|fp_reg_kernel|pass13_kernel|pass2_kernel|prog_kernel" -c 8 \
  -f -o gpurun_out/s7_full python tools/prof_one.py 16384 16384 all 1 > gpurun_out/s7_ncu.log 2>&1
ncu -i gpurun_out/s7_full.ncu-rep --page raw --csv > gpurun_out/s7_full_raw.csv 2>/dev/null
python tools/ncu_traffic.py gpurun_out/s7_full_raw.csv profiles/ncu_traffic_r01.json gpurun_out/s7_ncu_full_summary.txt "round 1 v9 kernels (tools/prof_one.py 16384 16384 all 1)" > /dev/null
cp profiles/ncu_traffic_r01.json gpurun_out/
for k in rowfft_kernel fp_reg_kernel pass2_kernel; do
  ncu -i gpurun_out/s7_full.ncu-rep --page source --csv --kernel-name regex:$k --print-source sass > gpurun_out/s7_src_$k.csv 2>/dev/null
done
rm -f gpurun_out/s7_full.ncu-rep
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/s7_bench.json 2> gpurun_out/s7_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s7_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/s7_launch_bench.log 2>&1
python tools/ncu_summary.py gpurun_out/s7_launches.csv > gpurun_out/s7_launches_summary.txt 2>&1
ls -la gpurun_out
