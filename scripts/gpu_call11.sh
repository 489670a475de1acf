#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tensor.json 2> gpurun_out/bench_tensor.err; echo "rc=$?"; tail -2 gpurun_out/bench_tensor.err
python -c "import json;d=json.load(open('gpurun_out/bench_tensor.json'));print('qps', round(d['value']), 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'], {k:round(v,3) for k,v in d['stages_ms_per_step'].items()}, d['cpu_baseline'].get('exact_path_ids_equal_to_cpu'), d['cpu_baseline'].get('value'))"
echo "== eval bench"; timeout 600 python scripts/eval_bench.py > gpurun_out/eval_bench.json 2> gpurun_out/eval_bench.err; echo "rc=$?"; tail -2 gpurun_out/eval_bench.err; cat gpurun_out/eval_bench.json
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -3
