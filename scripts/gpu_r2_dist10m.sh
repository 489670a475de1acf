#!/bin/bash
# configs[2]'s corpus (10M x 128, batch 1024 per GPU, ef_search 400) over 8 GPUs in the default multi-GPU mode (distributed scoring)
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --n-items 10000000 --batch 1024 --ef 400 --steps 5 --warmup 3 > gpurun_out/r2_bench_dist_10m_n8.json 2> gpurun_out/r2_bench_dist_10m_n8.err; echo "bench rc=$?"
grep "\[bench\]\|Error" gpurun_out/r2_bench_dist_10m_n8.err | grep -v "rank [1-9]" | tail -6
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_dist_10m_n8.json"))
print({k: d.get(k) for k in ("metric", "value", "ms_per_step", "gpu_launches", "recall_at_k_vs_bruteforce", "recall_target", "recall_held", "ids_bit_identical_to_one_gpu_search", "rows_scored_per_query")})
print(d["e2e"]["value"], d["roofline"]["frac"], d["stages_ms_per_step"], d["clocks"]); print(d.get("replica_mode"))
PY
