"""Throughput of the `main.py --job-type test` traversal (nann_search_eval_batch) on the 1M bench corpus, next to the
CPU oracle on the host cores (one query per core).  python scripts/eval_bench.py > gpurun_out/eval_bench.json"""
import json, os, sys, time
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nann_b200 as nb
from nann_b200 import index as nix, scorer_weights as sw
from oracle import oracle as orc
import bench

sh = bench.get_shard(1_000_000, 1, 0, "cuda:0")
ix = nb.Index.from_arrays(sh["emb"], sh["item_ids"], sh["ep"], sh["values"], sh["row_splits"])
W = sw.mlp_weights(seed=3)
sc = nb.Scorer.mlp(*W)
B, NS, TK, K = 256, (3, 1, 1), (400, 200, 100), 200
es = nb.EvalSearcher(ix, sc, B, TK, K)
queries = nix.synthetic_queries(sh["emb"], B * 8, seed=2)
out = {"workload": "1M x 128 corpus, batch 256, num_scoring_per_level [3,1,1], top_k_per_level [400,200,100], topk_eval 200 (reference defaults)"}
exact = es.search(queries[:B], NS, TK, K)
for prec, name in ((nb.SCORER_EXACT, "exact"), (nb.SCORER_TENSOR, "tensor")):
    sc.set_precision(prec)
    es.search(queries[:B], NS, TK, K)
    torch.cuda.synchronize(); t = time.perf_counter()
    rows = 0
    for i in range(1, 8):
        r = es.search(queries[i * B:(i + 1) * B], NS, TK, K)
        rows += r["n_scored"]
    dt = time.perf_counter() - t
    out[name] = {"queries_per_s": 7 * B / dt, "rows_scored_per_query": rows / (7 * B), "failed": int((r["status"] != 0).sum())}
# CPU oracle, one query per core, bounded sample
oix = orc.Index(sh["emb"], sh["item_ids"], sh["ep"].astype(np.int32), [v.astype(np.int32) for v in sh["values"]], sh["row_splits"])
om = orc.Mlp(*W)
cores = os.cpu_count() or 1
nq = min(B, 4 * cores)
def one(q):
    u = queries[q]
    return oix.search_eval(lambda rnd, ids: om.score(u, sh["emb"], ids) if len(ids) else np.zeros(0, np.float32), NS, TK, K)
t = time.perf_counter()
with ThreadPoolExecutor(cores) as ex:
    res = list(ex.map(one, range(nq)))
dt = time.perf_counter() - t
out["cpu_oracle"] = {"queries_per_s": nq / dt, "cores": cores, "sample": nq}
out["exact_ids_equal_to_cpu"] = bool(all(np.array_equal(exact["ids"][q], res[q]["ids"]) for q in range(nq)))
print(json.dumps(out))
