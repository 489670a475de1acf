#!/bin/bash
# round-2 call B (1 GPU): shard-group + builder tests first, then the rest, then builder timing and the default bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_search_gpu.py -m gpu -x -q -k "shard_group" 2>&1 | tail -8
timeout 900 python -m pytest tests/test_builder_gpu.py -m gpu -q 2>&1 | tail -25
timeout 600 python - <<'PY' 2>&1 | tail -12
import time, numpy as np, torch
from nann_b200 import builder, index as nix
for n in (1_000_000,):
    emb = nix.synthetic_corpus(n, 128, seed=0)
    e = torch.from_numpy(emb).cuda()
    for rep in range(2):
        t = time.time(); g = builder.build_hnsw(e, M=32, seed=4, return_stats=True); dt = time.time() - t
        print(f"builder n={n}: {dt:.2f}s", g["stats"], [len(v) for v in g["values"]], flush=True)
    # candidate quality: exact 64-NN of 2000 sample rows vs level-0 rows
    idx = np.random.default_rng(0).integers(0, n, 2000)
    d2 = 2 - 2 * (e[idx] @ e.T)
    d2[torch.arange(2000), torch.from_numpy(idx).cuda()] = float("inf")
    nn = torch.topk(d2, 32, dim=1, largest=False).indices.cpu().numpy()
    rs, v = g["row_splits"][0], g["values"][0]
    hit = np.mean([len(set(nn[q, :8].tolist()) & set(v[rs[i]:rs[i + 1]].tolist())) / 8 for q, i in enumerate(idx)])
    print("fraction of the 8 exact nearest neighbours present in the level-0 row:", hit, "mean degree", len(v) / n)
PY
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_builder_gpu.py 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
tail -4 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_n1.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "recall_at_k_vs_bruteforce", "rows_scored_per_query")})
print(d["e2e"]["value"], d["roofline"]["frac"], d["stages_ms_per_step"], d["clocks"])
print(d["cpu_baseline"])
PY
