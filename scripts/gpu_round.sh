#!/bin/bash
# One gpurun call: TC check (both kernels, separate processes), GPU tests, benches, ncu captures.
set -u
mkdir -p gpurun_out
echo "== tc check v2"; NANN_TC_KERNEL=2 timeout 300 python scripts/tc_check.py 2>&1 | tail -4 | tee gpurun_out/tc_check_v2.log
echo "== tc check v1"; NANN_TC_KERNEL=1 timeout 300 python scripts/tc_check.py 2>&1 | tail -4 | tee gpurun_out/tc_check_v1.log
if grep -q "PASS" gpurun_out/tc_check_v2.log; then TCV=2; else TCV=1; fi
export NANN_TC_KERNEL=$TCV; echo "using tc kernel v$TCV"
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench tensor (default)"; timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_tensor.json 2> gpurun_out/bench_tensor.err; echo "rc=$?"; tail -2 gpurun_out/bench_tensor.err; cut -c1-300 gpurun_out/bench_tensor.json
echo "== bench tensor v1 kernel"; NANN_TC_KERNEL=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --eval-queries 0 > gpurun_out/bench_tensor_v1.json 2>/dev/null; cut -c1-200 gpurun_out/bench_tensor_v1.json
echo "== bench exact"; timeout 900 python bench.py --steps 5 --warmup 3 --precision exact --no-cpu-baseline --eval-queries 0 > gpurun_out/bench_exact.json 2>/dev/null; cut -c1-200 gpurun_out/bench_exact.json
echo "== ncu launch list (tensor)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base mangled -k regex:nann -c 200 --csv \
   --log-file gpurun_out/launches_tensor.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eval-queries 0 > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full (tc scorer)"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:mlp_tc -s 7 -c 1 \
   -o gpurun_out/prof_mlp_tc -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eval-queries 0 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
echo "== ncu full (expand_filter)"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:expand_filter -s 5 -c 1 \
   -o gpurun_out/prof_expand -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eval-queries 0 > gpurun_out/ncu_full2.log 2>&1; echo "rc=$?"
ls -la gpurun_out | tail -25
