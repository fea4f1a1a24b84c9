#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/s9_pytest.txt
for v in 0 1; do
  echo "== VPFP_ROWFFT4=$v" >> gpurun_out/s9_rowfft.txt
  VPFP_ROWFFT4=$v timeout 300 python tools/time_ops.py 16384 16384 "edfdv_exp(table)" 2>&1 | tail -2 >> gpurun_out/s9_rowfft.txt
done
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"rowfft4_kernel" -c 1 \
  -f -o gpurun_out/s9_full python tools/prof_one.py 16384 16384 edfdv 1 > gpurun_out/s9_ncu.log 2>&1
ncu -i gpurun_out/s9_full.ncu-rep --page raw --csv > gpurun_out/s9_full_raw.csv 2>/dev/null
ncu -i gpurun_out/s9_full.ncu-rep --page source --csv --print-source sass > gpurun_out/s9_rowfft4_src.csv 2>/dev/null
rm -f gpurun_out/s9_full.ncu-rep
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err
ls -la gpurun_out
