#!/bin/bash
# usage: gpu_callN.sh N [extra bench args]   -- the driver's multi-GPU launch line
set -u
N=$1; shift
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo rc=$?
grep "bench\]" gpurun_out/bench_${N}gpu.err | tail -$N; wc -l gpurun_out/bench_${N}gpu.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${N}gpu.json').read())
print('qps', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'p50', round(d['latency_ms']['p50'],2), 'stages', {k:round(x,2) for k,x in d['stages_ms_per_step'].items()}, 'recall', d.get('recall_at_k_vs_bruteforce'), d.get('recall_error'), d['clocks'], d['config']['shard_level_topn'], 'rows/q', round(d['rows_scored_per_query']))
PY
