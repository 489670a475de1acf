#!/bin/bash
set -u
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo rc=$?
tail -3 gpurun_out/bench_2gpu.err; wc -l gpurun_out/bench_2gpu.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_2gpu.json').read())
print('qps', round(d['value']), 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'stages', {k:round(x,2) for k,x in d['stages_ms_per_step'].items()}, 'recall', d.get('recall_at_k_vs_bruteforce'), d.get('recall_error'), d['clocks'], d['config']['shard_level_topn'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err; echo rc=$?; cut -c1-200 gpurun_out/bench_2gpu_ref.json
