#!/bin/bash
# One gpurun call: GPU tests, bench (exact), ncu launch list + full capture, tensor-core check, bench (tensor).
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench exact"; timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_exact.json 2> gpurun_out/bench_exact.err; echo "rc=$?"; tail -2 gpurun_out/bench_exact.err; cut -c1-400 gpurun_out/bench_exact.json
echo "== ncu launch list (exact)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base mangled -k regex:nann -c 200 --csv \
   --log-file gpurun_out/launches_exact.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eval-queries 0 > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full (exact scorer, level-0 round)"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:mlp_exact -s 7 -c 1 \
   -o gpurun_out/prof_mlp_exact -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eval-queries 0 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
echo "== ncu full (expand_filter + topk)"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:"expand_filter|topk_kernel" -s 9 -c 2 \
   -o gpurun_out/prof_traverse -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eval-queries 0 > gpurun_out/ncu_full2.log 2>&1; echo "rc=$?"
echo "== tensor-core scorer check"; timeout 600 python scripts/tc_check.py 2>&1 | tail -15 | tee gpurun_out/tc_check.log
if grep -q "PASS" gpurun_out/tc_check.log; then
  echo "== bench tensor"; timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 --precision tensor > gpurun_out/bench_tensor.json 2> gpurun_out/bench_tensor.err; echo "rc=$?"; tail -2 gpurun_out/bench_tensor.err; cut -c1-400 gpurun_out/bench_tensor.json
fi
ls -la gpurun_out | tail -15
