"""GPU check of the tensor-core scorer against the exact scorer (run under gpurun)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import nann_b200 as nb
from nann_b200 import index as nix, scorer_weights as sw

emb = nix.synthetic_corpus(200000, 128, seed=0)
W = sw.mlp_weights()
rng = np.random.default_rng(0)
se = nb.Scorer.mlp(*W)
st = nb.Scorer.mlp(*W)
st.set_precision(nb.SCORER_TENSOR)
emb_d = torch.from_numpy(emb).cuda()
worst = 0.0
for n in (2, 127, 128, 129, 1000, 14800, 100000):
    ids = rng.integers(0, emb.shape[0], n).astype(np.int32)
    u = nix.synthetic_queries(emb, 1, seed=n)[0]
    a = nb.score_ids(se, u, emb_d, ids)
    b = nb.score_ids(st, u, emb_d, ids)
    d = np.abs(a - b)
    worst = max(worst, float(d.max()))
    print(f"n={n:7d} max|tc-exact|={d.max():.3e} mean={d.mean():.3e} |score| max={np.abs(a).max():.3f} nan={np.isnan(b).sum()}", flush=True)
print("WORST", worst, "PASS" if worst <= 1e-5 else "FAIL")
ids = rng.integers(0, emb.shape[0], 4_000_000).astype(np.int32)
ids_d = torch.from_numpy(ids).cuda()
u = nix.synthetic_queries(emb, 1, seed=1)[0]
for name, s in (("exact", se), ("tensor", st)):
    nb.score_ids(s, u, emb_d, ids_d[:200000])
    torch.cuda.synchronize(); t = time.perf_counter()
    out = nb.score_ids(s, u, emb_d, ids_d)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"{name}: {len(ids)/dt/1e6:.1f} M rows/s  {len(ids)*787456/dt/1e12:.1f} TFLOP/s (algorithmic, incl. op-call overhead)", flush=True)
