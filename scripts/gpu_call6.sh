#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== bench default"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tensor.json 2> gpurun_out/bench_tensor.err; echo "rc=$?"; tail -2 gpurun_out/bench_tensor.err
python -c "import json;d=json.load(open('gpurun_out/bench_tensor.json'));print('qps', round(d['value']), 'e2e', round(d['e2e']['value']), d['clocks'], d['cpu_baseline'])"
echo "== configs[2]: 10M x 128, batch 1024, ef 400"; free -g | head -2; df -h /tmp | tail -1
timeout 1500 python bench.py --n-items 10000000 --batch 1024 --ef 400 --steps 5 --warmup 3 --eval-queries 8 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "rc=$?"; tail -5 gpurun_out/bench_c3.err
python -c "import json;d=json.load(open('gpurun_out/bench_c3.json'));print('qps', round(d['value']), 'e2e', round(d['e2e']['value']), d['clocks'], d['stages_ms_per_step'], d['rows_scored_per_query'], d['roofline']['frac'], d.get('recall_at_k_vs_bruteforce'), d['cpu_baseline'])"
