#!/bin/bash
# round-2 call F (1 GPU): new tests; ncu captures: knn_filter_kernel (builder), bench launch list, mlp_tc8_kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_search_gpu.py tests/test_parity_baseline_scale.py -m gpu -x -q -k "graph or configs1" 2>&1 | tail -6
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_filter_kernel --launch-skip 10 -c 1 -o gpurun_out/r2_ncu_knn -f python scripts/builder_bench.py 1000000 1 > gpurun_out/r2_ncu_knn.log 2>&1; echo "ncu knn rc=$?"
ncu -i gpurun_out/r2_ncu_knn.ncu-rep --page raw --csv > gpurun_out/r2_ncu_knn_raw.csv 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --eval-queries 0 --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_tc8_kernel --launch-skip 12 -c 1 -o gpurun_out/r2_ncu_tc8 -f python bench.py --steps 2 --warmup 3 --eval-queries 0 --no-cpu-baseline > gpurun_out/r2_ncu_tc8.log 2>&1; echo "ncu tc8 rc=$?"
python scripts/ncu_summary.py gpurun_out/r2_ncu_summary.json gpurun_out/r2_ncu_knn.ncu-rep gpurun_out/r2_ncu_tc8.ncu-rep > /dev/null 2>&1; echo "summary rc=$?"
ls -la gpurun_out | tail -12
