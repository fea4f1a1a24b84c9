#!/bin/bash
# A/B of library builds on N GPUs: the product library and every vlapy_b200/lib/libvpfp_b200_<name>.so candidate,
# each through `bench.py --gpus N --no-e2e --no-cpu` (device-timed step, per-kernel times, parity against the single-GPU path)
#   tools/ab_multi.sh TAG N [workload]
TAG=$1; N=$2; W=${3:-c5}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/${TAG}_ab_${N}gpu.txt
: > $O
run() {
  echo "== $1" >> $O
  VPFP_B200_LIB=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29733 \
    bench.py --gpus $N --workload $W --steps ${STEPS:-20} --warmup 3 --no-e2e --no-cpu 2> gpurun_out/${TAG}_ab.err | python -c "
import json,sys
for l in sys.stdin:
    if not l.startswith('{'): continue
    d=json.loads(l); r=d['roofline']
    print('ms_per_step', d['ms_per_step'], 'parity_f', d['parity'].get('max_rel_err_f_vs_single'))
    print('  operators', {k: round(v,4) for k,v in r['operators_ms_per_step'].items()})
    print('  kernels', {k: round(v['ms_per_launch'],4) for k,v in r['kernels'].items()})
" >> $O
  tail -n 3 gpurun_out/${TAG}_ab.err | cut -c1-300 >> $O
}
run product ""
for so in vlapy_b200/lib/libvpfp_b200_*.so; do
  [ -f "$so" ] || continue
  run $so $PWD/$so
done
cat $O
