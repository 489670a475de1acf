#!/bin/bash
# BASELINE configs[3]: 100M x 128 row-sharded over 8 B200 (12.5M rows + their own HNSW per GPU)
mkdir -p gpurun_out
free -g | head -2
GB=$(free -g | awk '/^Mem:/{print $2}')
if [ "$GB" -lt 160 ]; then echo "host has only ${GB} GB of RAM: 8 x 12.5M-row shards need ~90 GB; not starting"; exit 0; fi
timeout 2400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --n-items 100000000 --steps 10 --warmup 3 --shard-scales 1.0837,1.275,1.5,1.5,0.2953 --cpu-shard-sample 32 > gpurun_out/r2_bench_100m.json 2> gpurun_out/r2_bench_100m.err; echo "bench rc=$?"
grep "\[bench\]" gpurun_out/r2_bench_100m.err | tail -12
tail -4 gpurun_out/r2_bench_100m.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_100m.json"))
print({k: d.get(k) for k in ("metric", "value", "ms_per_step", "gpu_launches", "recall_at_k_vs_bruteforce", "recall_target", "recall_held", "shard_beam_scale", "rows_scored_per_query", "exchange_exposed_ms_per_step")})
print(d["e2e"]["value"], d["roofline"]["frac"], d["stages_ms_per_step"], d["clocks"])
print(d.get("cpu_port_one_shard")); print(d["config"])
PY
