#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --n-items 10000000 --batch 1024 --ef 400 --steps 5 --warmup 3 > gpurun_out/r2_bench_configs2.json 2> gpurun_out/r2_bench_configs2.err; echo rc=$?
grep "\[bench\]" gpurun_out/r2_bench_configs2.err | tail -3
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_configs2.json"))
print({k: d[k] for k in ("value", "ms_per_step", "recall_at_k_vs_bruteforce", "rows_scored_per_query")})
print(d["e2e"]["value"], d["roofline"]["frac"], d["stages_ms_per_step"], d["clocks"]); print(d["cpu_baseline"])
PY
