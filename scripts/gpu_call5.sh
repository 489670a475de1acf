#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== bench tensor (full line)"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_tensor.json 2> gpurun_out/bench_tensor.err; echo "rc=$?"; tail -2 gpurun_out/bench_tensor.err; cut -c1-400 gpurun_out/bench_tensor.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_reference.json
echo "== sweep (configs[4])"; DUR=3 timeout 600 python scripts/sweep.py 2>&1 | tail -12
echo "== ncu launch list (tensor)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base mangled -k regex:nann -c 240 --csv \
   --log-file gpurun_out/launches_tensor.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eval-queries 0 > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "== ncu full (tc scorer)"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:mlp_tc8 -s 7 -c 1 \
   -o gpurun_out/prof_mlp_tc8 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eval-queries 0 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out | tail -12
