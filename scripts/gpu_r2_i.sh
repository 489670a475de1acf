#!/bin/bash
# 2 GPUs: IPC test, N=2 bench with the per-beam calibration, then the N=1 bench and the reference arm on the same box
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_shard_group_2gpu.py -m gpu -x -q 2>&1 | tail -3
bash scripts/gpu_r2_scale.sh 2 2>&1 | tail -9
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$?"; cat gpurun_out/r2_bench_reference.json | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_n1.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "recall_at_k_vs_bruteforce", "rows_scored_per_query")})
print(d["e2e"]["value"], d["roofline"]["frac"], d["stages_ms_per_step"], d["clocks"], d["latency_ms"])
print(d["cpu_baseline"]); print(d["roofline_gather"]["frac"] if d.get("roofline_gather") else None)
PY
