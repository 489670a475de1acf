#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== mma_ring"; timeout 120 scripts/micro/mma_ring 2>&1 | tee gpurun_out/mma_ring.txt
echo "== long bench (300 steps)"
timeout 600 python bench.py --steps 300 --warmup 3 --no-cpu-baseline --eval-queries 0 > gpurun_out/bench_long.json 2> gpurun_out/bench_long.err
python -c "import json;d=json.load(open('gpurun_out/bench_long.json'));print('long qps', round(d['value']), 'e2e', round(d['e2e']['value']), d['clocks'], 'score_ms', d['stages_ms_per_step']['score'])"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --eval-queries 0 > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err
python -c "import json;d=json.load(open('gpurun_out/bench_short.json'));print('short qps', round(d['value']), 'e2e', round(d['e2e']['value']), d['clocks'], 'score_ms', d['stages_ms_per_step']['score'])"
