#!/bin/bash
# round-2 call A (1 GPU): gpu tests, then the default bench (configs[1]) for both arms
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -5 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_n1.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "recall_at_k_vs_bruteforce", "rows_scored_per_query")})
print(d["e2e"]["value"], d["roofline"]["frac"], d["stages_ms_per_step"], d["clocks"])
print(d["cpu_baseline"])
PY
