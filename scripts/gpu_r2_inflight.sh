#!/bin/bash
# final check of the build (suite, smoke, default bench), then the same bench with 2 and 3 batches in flight
mkdir -p gpurun_out
bash scripts/gpu_r2_final.sh
for p in 2 3; do
  timeout 300 python bench.py --steps 20 --warmup 4 --inflight $p --no-cpu-baseline > gpurun_out/r2_bench_n1_inflight$p.json 2> gpurun_out/r2_bench_n1_inflight$p.err; echo "inflight $p rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_n1_inflight$p.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "batches_in_flight", "recall_at_k_vs_bruteforce")}, d["e2e"]["value"], d["latency_ms"]["p50"], d["clocks"])
PY
done
timeout 300 python bench.py --steps 20 --warmup 4 --inflight 1 --no-cpu-baseline > gpurun_out/r2_bench_n1_inflight1.json 2> gpurun_out/r2_bench_n1_inflight1.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_n1_inflight1.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "batches_in_flight")}, d["e2e"]["value"], d["latency_ms"]["p50"])
PY
