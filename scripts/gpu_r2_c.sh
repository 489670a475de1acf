#!/bin/bash
# round-2 call C (2 GPUs): the IPC shard-group test, then bench at N=2 (calibrated beams, in-library exchange, replica mode)
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8
timeout 600 python -m pytest tests/test_shard_group_2gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench rc=$?"
grep "\[bench\]" gpurun_out/r2_bench_n2.err | tail -12
tail -3 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_n2.json"))
print({k: d.get(k) for k in ("metric", "value", "ms_per_step", "gpu_launches", "recall_at_k_vs_bruteforce", "recall_target", "recall_held", "shard_beam_scale", "rows_scored_per_query", "exchange")})
print(d["e2e"]["value"], d["roofline"]["frac"], d["stages_ms_per_step"], d["clocks"])
print(d.get("calibration")); print(d.get("replica_mode")); print(d.get("nvlink")); print(d["config"]["shard_level_topn"])
PY
