#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -x 2>&1 | tail -4
NANN_TC_KERNEL=8 timeout 300 python scripts/tc_timeline.py 2>&1 | grep -v "ring stages" > gpurun_out/tc_timeline_v8.log; tail -3 gpurun_out/tc_timeline_v8.log
for i in 1 2; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --eval-queries 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('qps', round(d['value']), 'e2e', round(d['e2e']['value']), {k:round(v,3) for k,v in d['stages_ms_per_step'].items()}, d['clocks']['sm_mhz'], round(d['roofline']['frac'],3))"; done
timeout 300 python bench.py --steps 300 --warmup 3 --no-cpu-baseline --eval-queries 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('long qps', round(d['value']), {k:round(v,3) for k,v in d['stages_ms_per_step'].items()}, d['clocks']['sm_mhz'])"
