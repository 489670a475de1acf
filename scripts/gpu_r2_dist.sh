#!/bin/bash
# distributed scoring at N GPUs: tests first (N=2 only), then the bench in --mode dist
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_search_gpu.py -m gpu -x -q -k "distributed_scoring" 2>&1 | tail -5
  timeout 600 python -m pytest tests/test_dist_group_2gpu.py -m gpu -x -q 2>&1 | tail -3
fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_dist_n$N.json 2> gpurun_out/r2_bench_dist_n$N.err; echo "bench rc=$?"
grep "\[bench\]\|Error\|error" gpurun_out/r2_bench_dist_n$N.err | grep -v "rank [1-9]" | tail -8
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_dist_n$N.json"))
print({k: d.get(k) for k in ("metric", "value", "ms_per_step", "gpu_launches", "recall_at_k_vs_bruteforce", "recall_target", "recall_held", "ids_bit_identical_to_one_gpu_search", "rows_scored_per_query", "one_gpu_rows_scored_per_query", "profile_pass_ms_per_step")})
print(d["e2e"]["value"], d["roofline"]["frac"], d["stages_ms_per_step"], d["clocks"], d["latency_ms"])
PY
