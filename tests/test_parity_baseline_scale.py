"""Parity at BASELINE scale (configs[1]: 1M x 128 f32 corpus, HNSW M=32, batch 256, ef_search 200 -> level_topn
[100,200,200,200,200,200]) -- the workload bench.py times, not the 6k-row corpora of the other test files.

  * EXACT scorer: item ids, ranks and scores of 256 queries bit-identical to the oracle (the CPU restatement of the
    reference's TF custom-op path) on the same index files;
  * TENSOR scorer (the benchmarked path): every traced score within 1e-5 of the exact definition and the integer
    traversal bit-exact GIVEN those scores, on 16 queries;
  * size-independent properties on the full batch: sorted scores, no duplicate ids, ids from the corpus, determinism.
The index is built once per module by the CUDA builder (~1.5 s)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
T = [100, 200, 200, 200, 200, 200]


@pytest.fixture(scope="module")
def big(oracle):
    import nann_b200 as nb
    from nann_b200 import builder, index as nix, scorer_weights as sw
    n = 1_000_000
    emb = nix.synthetic_corpus(n, 128, seed=0)
    ids = nix.synthetic_item_ids(n, seed=1)
    g = builder.build_hnsw(emb, M=32, start_level=2, seed=4)
    W = sw.mlp_weights(seed=3)
    w = dict(nb=nb, emb=emb, ids=ids, g=g, W=W, queries=nix.synthetic_queries(emb, 256, seed=2))
    w["ix"] = nb.Index.from_arrays(emb, ids, g["enter_points"], g["values"], g["row_splits"])
    w["oix"] = oracle.Index(emb, ids, g["enter_points"].astype(np.int32), [v.astype(np.int32) for v in g["values"]], g["row_splits"])
    w["omlp"] = oracle.Mlp(*W)
    return w


def test_exact_path_bit_equal_on_256_queries_of_configs1(big):
    nb = big["nb"]
    sc = nb.Scorer.mlp(*big["W"])
    got = nb.Searcher(big["ix"], sc, 256, T).search(big["queries"], T)
    want = big["oix"].search_batch_mlp(big["omlp"], big["queries"], T, nthreads=os.cpu_count() or 1)
    assert np.all(got["status"] == 0) and np.all(want["status"] == 0)
    np.testing.assert_array_equal(got["ids"], want["ids"])
    np.testing.assert_array_equal(got["scores"].view(np.uint32), want["scores"].view(np.uint32))
    assert got["n_scored"].sum() == want["n_scored"]


def test_tensor_path_scores_and_traversal_on_configs1(big):
    nb = big["nb"]
    sc = nb.Scorer.mlp(*big["W"])
    sc.set_precision(nb.SCORER_TENSOR)
    se = nb.Searcher(big["ix"], sc, 256, T)
    full = se.search(big["queries"], T)
    again = se.search(big["queries"], T)
    assert np.all(full["status"] == 0)
    np.testing.assert_array_equal(full["ids"], again["ids"])                          # deterministic
    np.testing.assert_array_equal(full["scores"].view(np.uint32), again["scores"].view(np.uint32))
    assert np.all(np.diff(full["scores"], axis=1) <= 0)                               # sorted, value descending
    assert all(len(set(r.tolist())) == T[5] for r in full["ids"])                     # no duplicates
    assert full["ids"].min() >= 0 and full["ids"].max() < 1_000_000
    s16 = nb.Searcher(big["ix"], sc, 16, T)
    s16.set_trace(True)
    users = big["queries"][:16]
    got = s16.search(users, T)
    np.testing.assert_array_equal(got["ids"], full["ids"][:16])                       # batch composition does not matter
    worst = 0.0
    for q in range(16):
        traced = [s16.trace(q, r) for r in range(5)]
        for r in range(5):
            want = big["omlp"].score(users[q], big["emb"], traced[r][0])
            worst = max(worst, float(np.abs(want - traced[r][1]).max()))
        ref = big["oix"].search(lambda r, ids, t=traced: t[r][1], T)
        np.testing.assert_array_equal(got["ids"][q], ref["ids"])
        np.testing.assert_array_equal(got["scores"][q].view(np.uint32), ref["scores"].view(np.uint32))
    assert worst <= 1e-5, worst
