#!/bin/bash
# One gpurun call: TC variants (separate processes), GPU tests, benches, ncu captures.
set -u
mkdir -p gpurun_out
echo "== tc checks"; bash scripts/tc_pick.sh > gpurun_out/tc_pick.txt; cat gpurun_out/tc_pick.txt
echo "== bench per tc version (short)"
best=1; best_q=0
for v in ${TC_VERSIONS:-1 2 3 4}; do
  if grep -q PASS gpurun_out/tc_check_v$v.log; then
    NANN_TC_KERNEL=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --eval-queries 0 > gpurun_out/bench_tc_v$v.json 2> gpurun_out/bench_tc_v$v.err
    q=$(python -c "import json;d=json.load(open('gpurun_out/bench_tc_v$v.json'));print(round(d['value']), round(d['stages_ms_per_step']['score'],2), round(d['stages_ms_per_step']['expand_filter'],2))" 2>/dev/null || echo "0 0 0")
    echo "v$v: qps score_ms expand_ms = $q"
    qq=$(echo $q | awk '{print $1}')
    if [ "$qq" -gt "$best_q" ]; then best=$v; best_q=$qq; fi
  fi
done
export NANN_TC_KERNEL=$best; echo "using tc kernel v$best ($best_q qps)"; echo $best > gpurun_out/tc_best.txt
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench tensor (full line)"; timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_tensor.json 2> gpurun_out/bench_tensor.err; echo "rc=$?"; tail -2 gpurun_out/bench_tensor.err; cut -c1-300 gpurun_out/bench_tensor.json
echo "== ncu launch list (tensor)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base mangled -k regex:nann -c 240 --csv \
   --log-file gpurun_out/launches_tensor.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eval-queries 0 > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full (tc scorer)"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:mlp_tc -s 7 -c 1 \
   -o gpurun_out/prof_mlp_tc -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eval-queries 0 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
echo "== ncu full (expand_filter)"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:expand_filter -s 5 -c 1 \
   -o gpurun_out/prof_expand -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eval-queries 0 > gpurun_out/ncu_full2.log 2>&1; echo "rc=$?"
ls -la gpurun_out | tail -25
