#!/bin/bash
# Where do the extra milliseconds of the N-GPU step go?  gpurun --gpus N -- bash scripts/gpu_scaling_attribution.sh N
N=${1:-2}
mkdir -p gpurun_out/r2/attr
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node=$N"
i=0
for flags in "" "--no-sync-bn" "--no-grad-allreduce" "--no-sync-bn --no-grad-allreduce" "--no-overlap"; do
  i=$((i+1))
  timeout 200 $R --master-port $((29900+i)) bench.py --gpus $N --steps 30 --warmup 5 --no-gpu-baseline $flags > gpurun_out/r2/attr/n${N}_$i.json 2> gpurun_out/r2/attr/n${N}_$i.err
  python -c "
import json;d=json.load(open('gpurun_out/r2/attr/n${N}_$i.json'));print('N=$N %-42s %8.0f img/s %6.2f ms/step  e2e %8.0f' % ('$flags' or '(full: SyncBN + overlapped grad all-reduce)', d['value'], d['ms_per_step'], d['e2e']['value']))"
done
timeout 200 python bench.py --steps 30 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r2/attr/n1.json 2>/dev/null
python -c "
import json;d=json.load(open('gpurun_out/r2/attr/n1.json'));print('N=1 %-42s %8.0f img/s %6.2f ms/step  e2e %8.0f' % ('', d['value'], d['ms_per_step'], d['e2e']['value']))"
# NCCL channel count (CTAs that the gradient all-reduce takes away from the persistent conv kernels)
for ch in 2 4 8; do
  NCCL_MAX_NCHANNELS=$ch timeout 200 $R --master-port $((29950+ch)) bench.py --gpus $N --steps 30 --warmup 5 --no-gpu-baseline > gpurun_out/r2/attr/n${N}_ch$ch.json 2> gpurun_out/r2/attr/n${N}_ch$ch.err
  python -c "
import json;d=json.load(open('gpurun_out/r2/attr/n${N}_ch$ch.json'));print('N=$N %-42s %8.0f img/s %6.2f ms/step  e2e %8.0f' % ('NCCL_MAX_NCHANNELS=$ch', d['value'], d['ms_per_step'], d['e2e']['value']))"
done
