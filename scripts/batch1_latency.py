"""batch-1 latency of the plain public call (host arrays in and out, no stream argument: the reference's operating mode) on
the 1M x 128 corpus, TENSOR scorer: CUDA-graph replay on the searcher's own stream vs the eager launch sequence
(NANN_GRAPH_MAX_BATCH=0 in a second process would be the cleaner A/B; here eager = a searcher with tracing on)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import nann_b200 as nb
from nann_b200 import scorer_weights as sw

T = bench.EF_TOPN[200]
sh = bench.get_shard(1_000_000, 1, 0, "cuda:0")
ix = nb.Index.from_arrays(sh["emb"], sh["item_ids"], sh["ep"], sh["values"], sh["row_splits"])
sc = nb.Scorer.mlp(*sw.mlp_weights(seed=3))
sc.set_precision(nb.SCORER_TENSOR)
q = bench.nix().synthetic_queries(sh["emb"], 400, seed=2)


def run(se, n=300):
    for i in range(20):
        se.search(q[i:i + 1], T)
    ts = []
    for i in range(n):
        t0 = time.perf_counter()
        se.search(q[20 + i:21 + i], T)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return {"p50_ms": 1e3 * ts[len(ts) // 2], "p99_ms": 1e3 * ts[int(len(ts) * 0.99)]}


replay = run(nb.Searcher(ix, sc, 1, T))
eager_se = nb.Searcher(ix, sc, 1, T)
eager_se.set_trace(True)
eager = run(eager_se)
print(json.dumps({"what": "nann_search_batch, B=1, host users in / host ids+scores out, stream=NULL, 1M x 128, TENSOR scorer",
                  "graph_replay_on_own_stream": replay, "eager_with_trace_copies": eager}))
