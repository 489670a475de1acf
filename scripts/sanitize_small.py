"""Small end-to-end run for compute-sanitizer (memcheck): exec.pb traversal (CTA expand+filter, top-k), eval traversal,
exact + tensor-core scorers on a 3000-item corpus.   compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nann_b200 as nb
from nann_b200 import index as nix, scorer_weights as sw

n = 3000
emb = nix.synthetic_corpus(n, 128, seed=0)
g = nix.build_hnsw(emb, M=16, m_levels=6, n_cand=40, device="cpu")
ix = nb.Index.from_arrays(emb, nix.synthetic_item_ids(n), g["enter_points"], g["values"], g["row_splits"])
sc = nb.Scorer.mlp(*sw.mlp_weights())
users = nix.synthetic_queries(emb, 5)
T = [20, 40, 40, 40, 40, 40]
se = nb.Searcher(ix, sc, 5, T)
a = se.search(users, T)
ev = nb.EvalSearcher(ix, sc, 5, (40, 20, 10), 20).search(users, (3, 1, 1), (40, 20, 10), 20)
if os.environ.get("SANITIZE_TENSOR", "1") == "1":
    sc.set_precision(nb.SCORER_TENSOR)
    b = se.search(users, T)
    print("tensor ok", np.mean([len(set(x.tolist()) & set(y.tolist())) for x, y in zip(a["ids"], b["ids"])]))
print("ok", a["status"], ev["status"], ev["n"])
