#!/bin/bash
# bench at N GPUs (arg 1) on the default corpus: sharded (calibrated) + replica mode in one line
N=${1:-4}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench rc=$?"
grep "\[bench\]" gpurun_out/r2_bench_n$N.err | grep -v "rank [1-9]" | tail -8
tail -3 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_n$N.json"))
print({k: d.get(k) for k in ("metric", "value", "ms_per_step", "gpu_launches", "recall_at_k_vs_bruteforce", "recall_target", "recall_held", "shard_beam_scale", "rows_scored_per_query", "one_gpu_rows_scored_per_query", "exchange_exposed_ms_per_step")})
print(d["e2e"]["value"], d["roofline"]["frac"], d["stages_ms_per_step"], d["clocks"])
print(len(d.get("calibration", [])), "calibration trials; last:", d.get("calibration", [None])[-1])
print(d.get("replica_mode")); print(d.get("nvlink")); print(d["config"]["shard_level_topn"])
PY
