#!/bin/bash
# One parametrised GPU-box session (replaces the numbered one-off scripts of round 1):
#     gpurun --timeout T -- 'bash tools/gpu_session.sh TAG STAGE [STAGE ...]'
# Every stage writes its result to gpurun_out/TAG_<stage>.* ; copy what is to be kept into profiles/.
# Stages
#   tests            pytest -m gpu (whole suite), tail kept
#   tests:<expr>     pytest -m gpu -k <expr>
#   multi            tests/test_gpu_multi.py (needs >= 2 GPUs on the box)
#   bench:<w>        python bench.py --workload <w>            (1 GPU; w = c1..c5)
#   benchN:<n>:<w>   torchrun bench.py --gpus n --workload w   (n GPUs)
#   ops:<list>       tools/time_ops.py 16384 16384 <list>
#   ab:<list>        the same on the product library and on every vlapy_b200/lib/libvpfp_b200_<name>.so (candidates
#                    built with -D flags); AB_TESTS=<expr> also runs the parity tests on each candidate
#   launches         ncu launch list (gpu__time_duration) of a 2-step C5 bench
#   full:<regex>     ncu --set full --import-source on of the kernels matching <regex> (tools/prof_one.py)
#   sanitizer        compute-sanitizer racecheck + memcheck on the kernels with barrier-free prefetch / peer stores
#   fp64peak         tools/fp64_peak (DFMA chain microbenchmark)
#   wirebench        tools/wirebench.cu (2 GPUs: NVLink-bound kernel beside a local copy)
#   smoke            __graft_entry__.smoke()
TAG=${1:-s}; shift
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/$TAG
NG=$(nvidia-smi -L | wc -l)
for st in "$@"; do
  name=${st%%:*}; arg=${st#*:}; [ "$arg" = "$st" ] && arg=""
  echo "=== stage $st ($(date +%T))"
  case $name in
    tests)
      if [ -n "$arg" ]; then K=(-k "$arg"); else K=(); fi
      ( timeout 1500 python -m pytest tests -m gpu -x -q "${K[@]}" 2>&1 | tail -15 ) > ${O}_tests.txt; tail -3 ${O}_tests.txt ;;
    multi)
      ( timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -rs 2>&1 | tail -15 ) > ${O}_multi_${NG}gpu.txt; tail -3 ${O}_multi_${NG}gpu.txt ;;
    bench)
      timeout 900 python bench.py --workload $arg --steps ${STEPS:-20} --warmup 3 > ${O}_bench_${arg}_1gpu.json 2> ${O}_bench_${arg}_1gpu.err; cut -c1-400 ${O}_bench_${arg}_1gpu.json ;;
    benchN)
      n=${arg%%:*}; w=${arg#*:}
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29721 \
        bench.py --gpus $n --workload $w --steps ${STEPS:-20} --warmup 3 > ${O}_bench_${w}_${n}gpu.json 2> ${O}_bench_${w}_${n}gpu.err
      cut -c1-400 ${O}_bench_${w}_${n}gpu.json ;;
    ops)
      timeout 300 python tools/time_ops.py 16384 16384 "$arg" > ${O}_ops.txt 2>&1; cat ${O}_ops.txt ;;
    ab)
      # every vlapy_b200/lib/libvpfp_b200_<name>.so present (candidates built with -D flags) against the product library
      { echo "== product library"; timeout 120 python tools/time_ops.py 16384 16384 "$arg" 2>&1 | tail -8
        for so in vlapy_b200/lib/libvpfp_b200_*.so; do
          [ -f "$so" ] || continue
          echo "== $so"; VPFP_B200_LIB=$PWD/$so timeout 120 python tools/time_ops.py 16384 16384 "$arg" 2>&1 | tail -8
          if [ -n "$AB_TESTS" ]; then echo "-- parity tests on $so"; VPFP_B200_LIB=$PWD/$so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$AB_TESTS" 2>&1 | tail -2; fi
        done; } > ${O}_ab.txt; cat ${O}_ab.txt ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches.csv \
        python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > ${O}_launches.log 2>&1
      python tools/ncu_summary.py ${O}_launches.csv > ${O}_launches.txt 2>&1; head -30 ${O}_launches.txt ;;
    full)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$arg" -c 4 -f -o ${O}_full \
        python tools/prof_one.py 16384 16384 all 1 > ${O}_full.log 2>&1; tail -2 ${O}_full.log ;;
    sanitizer)
      SEL="test_fp_sizes_vs_oracle and 4096 or test_single_pass_row_kernel and 257 or test_vdfdx_fused_density and 4096 or test_vdfdx_fused_density_of_an_ensemble or test_tridiag or test_mid_size_single_pass_kernels and 512 or test_semi_lagrangian_operators or test_moments_many_short_rows and 512"
      ( timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" 2>&1 | tail -40 ) > ${O}_racecheck.txt
      ( timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" 2>&1 | tail -25 ) > ${O}_memcheck.txt
      tail -n 4 ${O}_racecheck.txt; tail -n 4 ${O}_memcheck.txt ;;
    wirebench)
      # needs 2 GPUs: a wire-bound peer-store kernel beside a local HBM copy, sharing SMs or confined to a few
      nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/wirebench.bin tools/wirebench.cu && timeout 120 ./tools/wirebench.bin > ${O}_wirebench.txt 2>&1; cat ${O}_wirebench.txt ;;
    fp64peak)
      timeout 120 python tools/fp64_peak.py > ${O}_fp64_peak.txt 2>&1; cat ${O}_fp64_peak.txt ;;
    smoke)
      ( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) > ${O}_smoke.txt; cat ${O}_smoke.txt ;;
    *) echo "unknown stage $st" ;;
  esac
done
ls -la gpurun_out | tail -20
