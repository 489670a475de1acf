#!/bin/bash
set -u
mkdir -p gpurun_out
export TC_VERSIONS="${TC_VERSIONS:-3 4}"
echo "== tc checks"; bash scripts/tc_pick.sh > gpurun_out/tc_pick.txt; cat gpurun_out/tc_pick.txt
for v in $TC_VERSIONS; do
  if grep -q PASS gpurun_out/tc_check_v$v.log; then
    echo "== timeline v$v"; NANN_TC_KERNEL=$v timeout 300 python scripts/tc_timeline.py 2>&1 | tail -64 | tee gpurun_out/tc_timeline_v$v.log | sed -n '/tile 11/,$p'
    NANN_TC_KERNEL=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --eval-queries 0 > gpurun_out/bench_tc_v$v.json 2> gpurun_out/bench_tc_v$v.err
    python -c "import json;d=json.load(open('gpurun_out/bench_tc_v$v.json'));print('v$v qps', round(d['value']), 'score_ms', round(d['stages_ms_per_step']['score'],2), 'roof', round(d['roofline']['achieved'],1))" 2>/dev/null || tail -3 gpurun_out/bench_tc_v$v.err
  fi
done
