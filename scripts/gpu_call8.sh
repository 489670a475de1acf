#!/bin/bash
set -u
mkdir -p gpurun_out
for v in 8 3; do
  NANN_BENCH_DEBUG=1 NANN_TC_KERNEL=$v python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$v bench.py --gpus 2 --steps 10 --warmup 3 --eval-queries 0 > gpurun_out/bench_2gpu_v$v.json 2> gpurun_out/bench_2gpu_v$v.err
  python - <<PY
import json
txt=open('gpurun_out/bench_2gpu_v$v.json').read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print('v$v', 'qps', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'stages', {k:round(x,2) for k,x in d['stages_ms_per_step'].items()}, d.get('debug'))
PY
done
