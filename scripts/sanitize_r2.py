"""Small end-to-end run of the round-2 code for compute-sanitizer (memcheck / initcheck): CUDA index builder (tcgen05
filter kernel, compaction, refinement, reverse links), BloomFilterDifference, the sharded exchange (two members, push /
back-pressure / wait / merge over three sequences), distributed scoring with a group of one, CUDA-graph replay.
   compute-sanitizer --tool memcheck python scripts/sanitize_r2.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
import torch
import nann_b200 as nb
from nann_b200 import builder, index as nix, scorer_weights as sw
from nann_b200.distributed import DistGroup, ShardGroup, shard_bounds, shard_level_topn

n = 1500
emb = nix.synthetic_corpus(n, 128, seed=0)
ids = nix.synthetic_item_ids(n)
g = builder.build_hnsw(emb, M=16, m_levels=6, seed=4, return_stats=True)
print("builder ok", g["stats"], [len(v) for v in g["values"]])
flags = np.zeros(8, np.int32)
c, crs, _ = nb.bloom_filter_difference(np.arange(300, dtype=np.int64) * 7919, np.array([0, 100, 300], np.int64), flags, bucket=0, bucket_size=8)
print("bloom ok", len(c), crs.tolist())
sc = nb.Scorer.mlp(*sw.mlp_weights())
T = [20, 40, 40, 40, 40, 40]
users = nix.synthetic_queries(emb, 12)
ix = nb.Index.from_arrays(emb, ids, g["enter_points"], g["values"], g["row_splits"])
side = torch.cuda.Stream()
se = nb.Searcher(ix, sc, 4, T)
a = [se.search(users[:2], T, stream=side) for _ in range(3)]          # eager, capture, replay
assert all(np.array_equal(a[0]["ids"], x["ids"]) for x in a)
print("graph replay ok")
# sharded exchange, two members on one device
Ts = shard_level_topn(T, 2)
members, searchers, keep = [ShardGroup(r, 2, 4, Ts[5]) for r in range(2)], [], []
ShardGroup.connect_local(members)
for r in range(2):
    lo, hi = shard_bounds(n, 2, r)
    gr = builder.build_hnsw(emb[lo:hi], M=16, m_levels=6, seed=4 + r)
    ixr = nb.Index.from_arrays(emb[lo:hi], ids[lo:hi], gr["enter_points"], gr["values"], gr["row_splits"])
    keep.append(ixr); searchers.append(nb.Searcher(ixr, sc, 4, Ts))
u_dev = torch.from_numpy(users).cuda()
outs = [[(torch.empty((4, T[5]), dtype=torch.int64, device="cuda"), torch.empty((4, T[5]), dtype=torch.float32, device="cuda"),
          torch.empty((4,), dtype=torch.int32, device="cuda")) for _ in range(3)] for _ in range(2)]
torch.cuda.synchronize()
for i in range(3):
    for r in range(2):
        members[r].push(searchers[r], u_dev[i * 4:(i + 1) * 4], Ts, stream=side)
    for r in range(2):
        members[r].merge(T[5], *outs[r][i])
for m in members:
    m.wait()
torch.cuda.synchronize()
assert torch.equal(outs[0][2][0], outs[1][2][0]) and int(outs[0][2][2].sum()) == 0
print("shard exchange ok")
# distributed scoring, group of one, both precisions
for prec in (nb.SCORER_EXACT, nb.SCORER_TENSOR):
    sc.set_precision(prec)
    ixd = nb.Index.from_arrays_sharded(n, emb, 0, ids, g["enter_points"], g["values"], g["row_splits"])
    sed = nb.Searcher(ixd, sc, 4, T)
    grp = DistGroup(sed, 0, 1)
    s_h, i_h, st_h = grp.search(users[:4], T)
    want = nb.Searcher(ix, sc, 4, T).search(users[:4], T)
    assert np.array_equal(i_h, want["ids"]) and np.array_equal(s_h.view(np.uint32), want["scores"].view(np.uint32))
    grp.close()
print("distributed scoring ok")
