#!/bin/bash
# Multi-GPU evidence in one gpurun call:  gpurun --gpus N -- bash scripts/gpu_multi_gpu_suite.sh N
#   DP equivalence (strict per-parameter gates) at N and N/2 ranks, bench.py scaling 1..N for the headline
#   config, the other BASELINE configs at N ranks.  Logs under gpurun_out/r2/mg/.
NMAX=${1:-8}
mkdir -p gpurun_out/r2/mg
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in $NMAX $((NMAX/2)); do
  [ $n -ge 2 ] || continue
  timeout 300 $R --nproc-per-node=$n --master-port $((29600+n)) tests/tools/dp_equivalence.py > gpurun_out/r2/mg/dp$n.log 2>&1; echo dp$n rc=$?
  grep "bottleneck\|resnet50 DP\|peer\|DP_EQ" gpurun_out/r2/mg/dp$n.log | cut -c1-330
done
n=$NMAX
while [ $n -ge 2 ]; do
  timeout 300 $R --nproc-per-node=$n --master-port $((29700+n)) bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/r2/mg/bench_r50_$n.json 2> gpurun_out/r2/mg/bench_r50_$n.err; echo r50 x$n rc=$?
  n=$((n/2))
done
timeout 200 python bench.py --steps 30 --no-cpu-baseline > gpurun_out/r2/mg/bench_r50_1.json 2> gpurun_out/r2/mg/bench_r50_1.err
for c in r50_128 bresnet50 r50_arcface; do
  timeout 300 $R --nproc-per-node=$NMAX --master-port 29811 bench.py --gpus $NMAX --config $c --steps 20 --warmup 5 --no-gpu-baseline > gpurun_out/r2/mg/bench_${c}_$NMAX.json 2> gpurun_out/r2/mg/bench_${c}_$NMAX.err; echo $c x$NMAX rc=$?
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2/mg/bench_*.json')):
    try:
        d=json.load(open(f)); gb=d.get('gpu_baseline') or {}
        print(f.split('/')[-1], d['n_gpus'], round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'graph', d['config']['cuda_graph'], 'gpu_baseline', round(gb.get('value',0)))
    except Exception as e: print(f, 'ERR', e)
PY
