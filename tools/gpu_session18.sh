#!/bin/bash
# one short session: parity of the two-CTAs-per-SM row kernel (rowfft2.cuh), its A/B timing against rowfft.cuh with
# the cache-hint / L2-prefetch knobs, and the bench line with either kernel as the library default
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_cta or (single_pass and 16384-257)" 2>&1 | tail -6 ) > gpurun_out/s18_pytest.txt
( timeout 100 python tools/time_row2.py 2>&1 | tail -9 ) > gpurun_out/s18_time.txt
( VPFP_ROWFFT2_HINTS=1 timeout 60 python tools/time_row2.py short 2>&1 | tail -3 ) >> gpurun_out/s18_time.txt
( VPFP_ROWFFT2_HINTS=3 timeout 60 python tools/time_row2.py short 2>&1 | tail -3 ) >> gpurun_out/s18_time.txt
( VPFP_ROWFFT_L2PF=1 timeout 60 python tools/time_row2.py short 2>&1 | tail -3 ) >> gpurun_out/s18_time.txt
( VPFP_ROWFFT_L2PF=1 VPFP_ROWFFT2_HINTS=1 timeout 60 python tools/time_row2.py short 2>&1 | tail -3 ) >> gpurun_out/s18_time.txt
cat gpurun_out/s18_pytest.txt gpurun_out/s18_time.txt
VPFP_ROWFFT2=1 timeout 240 python bench.py --steps 10 --warmup 3 > gpurun_out/s18_bench_two_cta.json 2> gpurun_out/s18_bench_two_cta.err
VPFP_ROWFFT2=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:"rowfft2_kernel" -c 1 \
  -f -o gpurun_out/s18_row2 python tools/prof_one.py 16384 16384 edfdv 1 > gpurun_out/s18_ncu.log 2>&1
ncu -i gpurun_out/s18_row2.ncu-rep --page raw --csv > gpurun_out/s18_row2_raw.csv 2>/dev/null
python tools/ncu_traffic.py gpurun_out/s18_row2_raw.csv gpurun_out/s18_row2_traffic.json gpurun_out/s18_row2_full.txt "round 1 v10 rowfft2 (tools/prof_one.py 16384 16384 edfdv 1, VPFP_ROWFFT2=1)" | head -30
VPFP_ROWFFT2=0 timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s18_bench_one_cta.json 2> gpurun_out/s18_bench_one_cta.err
cut -c1-400 gpurun_out/s18_bench_two_cta.json gpurun_out/s18_bench_one_cta.json
tail -3 gpurun_out/s18_bench_two_cta.err
