#!/bin/bash
# round-1 v14: two-stage density reduction -- parity subset, bench line
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "fused_density or vp50 or nlepw_c2 or vdfdx_fullsize or ensemble_matches" 2>&1 | tail -4 ) > gpurun_out/s25_pytest.txt
cat gpurun_out/s25_pytest.txt
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/s25_bench_1gpu.json 2> gpurun_out/s25_bench_1gpu.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/s25_bench_1gpu.json"))
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"], {k: round(v["ms_per_launch"], 4) for k, v in d["roofline"]["kernels"].items()})
PY
