"""Launch list of ONE distributed-scoring step (a group of one on one GPU: the whole exchange path runs -- bucket -> request
window -> scorer on the pseudo-queries -> return -> unbucket, flags and one-warp waits -- with the window in local HBM), for
`ncu --profile-from-start off --metrics gpu__time_duration.sum`: what the exchange kernels cost next to the scorer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import nann_b200 as nb
from nann_b200 import scorer_weights as sw
from nann_b200.distributed import DistGroup

T = bench.EF_TOPN[200]
full = bench.get_shard(1_000_000, 1, 0, "cuda:0")
n = full["emb"].shape[0]
ix = nb.Index.from_arrays_sharded(n, full["emb"], 0, full["item_ids"], full["ep"], full["values"], full["row_splits"])
sc = nb.Scorer.mlp(*sw.mlp_weights(seed=3))
sc.set_precision(nb.SCORER_TENSOR)
B = 256
se = nb.Searcher(ix, sc, B, T)
grp = DistGroup(se, 0, 1)
q = torch.from_numpy(bench.nix().synthetic_queries(full["emb"], 4 * B, seed=2)).cuda()
out = (torch.empty((B, T[5]), dtype=torch.int64, device="cuda"), torch.empty((B, T[5]), dtype=torch.float32, device="cuda"))
st = torch.cuda.Stream()
for i in range(3):
    grp.search(q[i * B:(i + 1) * B], T, *out, stream=st)
torch.cuda.synchronize()
torch.cuda.profiler.start()
grp.search(q[3 * B:4 * B], T, *out, stream=st)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
grp.check()
print("ok")
