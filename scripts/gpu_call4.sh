#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu (default tc kernel)"; timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== bench (default)"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tensor.json 2> gpurun_out/bench_tensor.err; echo "rc=$?"; tail -2 gpurun_out/bench_tensor.err
python -c "import json;d=json.load(open('gpurun_out/bench_tensor.json'));print('qps', round(d['value']), 'e2e', round(d['e2e']['value']), d['clocks'], d['stages_ms_per_step'], 'roof', d['roofline']['frac'], d['cpu_baseline'])"
echo "== long"; timeout 600 python bench.py --steps 300 --warmup 3 --no-cpu-baseline --eval-queries 0 > gpurun_out/bench_long.json 2> gpurun_out/bench_long.err
python -c "import json;d=json.load(open('gpurun_out/bench_long.json'));print('long qps', round(d['value']), 'e2e', round(d['e2e']['value']), d['clocks'], 'score_ms', d['stages_ms_per_step']['score'])"
