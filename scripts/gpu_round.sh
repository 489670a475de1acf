#!/bin/bash
# One gpurun call: GPU tests, smoke, bench lines (both arms), ncu launch list + full captures of the two main kernels.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
echo "== bench (this repo)"; timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_tensor.json 2> gpurun_out/bench_tensor.err; echo "rc=$?"; tail -2 gpurun_out/bench_tensor.err; cut -c1-300 gpurun_out/bench_tensor.json
echo "== bench (reference arm)"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"; cut -c1-200 gpurun_out/bench_reference.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base mangled -k regex:nann -c 400 --csv \
   --log-file gpurun_out/launches_tensor.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eval-queries 0 > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full (tensor-core scorer)"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:mlp_tc8 -s 7 -c 1 \
   -o gpurun_out/prof_mlp_tc8 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eval-queries 0 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
echo "== ncu full (expand_filter)"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:expand_filter -s 5 -c 1 \
   -o gpurun_out/prof_expand -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eval-queries 0 > gpurun_out/ncu_full2.log 2>&1; echo "rc=$?"
ls -la gpurun_out | tail -14
