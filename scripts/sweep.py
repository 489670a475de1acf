"""BASELINE configs[4]: dynamic-batch latency/throughput sweep 1..4096 queries on the 1M corpus
(blaze-benchmark parity): closed-loop executor, predictor_num searchers x max_batch_size."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import nann_b200 as nb
from nann_b200 import harness, index as nix, scorer_weights as sw
import bench

T = bench.EF_TOPN[200]
sh = bench.get_shard(int(os.environ.get("N_ITEMS", 1_000_000)), 1, 0, "cuda:0")
ix = nb.Index.from_arrays(sh["emb"], sh["item_ids"], sh["ep"], sh["values"], sh["row_splits"])
sc = nb.Scorer.mlp(*sw.mlp_weights(seed=3))
if os.environ.get("PRECISION", "tensor") == "tensor":
    sc.set_precision(nb.SCORER_TENSOR)
queries = nix.synthetic_queries(sh["emb"], 16384, seed=2)
out = []
for mb, pn in [(1, 1), (1, 4), (4, 4), (16, 4), (64, 4), (256, 2), (256, 4), (1024, 2), (4096, 1)]:
    r = harness.run_benchmark(ix, sc, T, queries, predictor_num=pn, bench_thread_count=4, duration=float(os.environ.get("DUR", 3)),
                              max_batch_size=mb)
    row = dict(max_batch_size=mb, predictor_num=pn, qps=r["throughput"], failures=r["failures"], batch_mean=r["batchsize"]["mean"],
               latency_us_median=r["latency_us"]["median"], latency_us_p99=r["latency_us"]["p99"],
               e2e_us_median=r["e2e_latency_us"]["median"], e2e_us_p99=r["e2e_latency_us"]["p99"])
    out.append(row)
    print(json.dumps(row), flush=True)
json.dump(dict(workload="1M items d=128, ef=200, MLP 2x512, closed loop (qps=-1)", precision=os.environ.get("PRECISION", "tensor"), rows=out),
          open("gpurun_out/sweep.json", "w"), indent=1)
