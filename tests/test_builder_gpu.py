"""The CUDA index builder (replaces the offline faiss step of build_hnsw_index.py:33-67).

Index construction has no reference-held vectors (faiss is a third-party dependency that is absent here and its
graph is not pinned by any reference test), so the builder is pinned two ways:
  * bit for bit against oracle.build_hnsw, the CPU statement of the same batch construction (exact fp32 arithmetic);
  * against the torch stand-in nann_b200/index.py (different arithmetic for pair distances): edge overlap.
and by what the search path needs from it: valid CSR, degree caps, no self links, rows closest-first."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check_structure(emb, g, M):
    n = emb.shape[0]
    for l, cap in ((0, 2 * M), (1, M)):
        v, rs = g["values"][l], g["row_splits"][l]
        assert rs.shape == (n + 1,) and rs[0] == 0 and rs[-1] == len(v) and np.all(np.diff(rs) >= 0)
        deg = np.diff(rs)
        assert deg.max() <= cap
        members = g["levels"] >= l
        assert np.all(deg[~members] == 0)
        if members.sum() > 1:
            assert np.all(deg[members] >= 1)
        src = np.repeat(np.arange(n), deg)
        assert np.all(v != src) and np.all(members[v])
        d = np.linalg.norm(emb[src].astype(np.float64) - emb[v].astype(np.float64), axis=1)
        same = src[1:] == src[:-1]
        assert np.all(d[1:][same] >= d[:-1][same] - 1e-6)           # closest first inside a row


@pytest.mark.parametrize("n,m_levels", [(5000, 6), (700, 4), (100, 3)])
def test_builder_bit_equal_to_cpu_statement(oracle, n, m_levels):
    from nann_b200 import builder, index as nix
    emb = nix.synthetic_corpus(n, 128, seed=11)
    M = 16
    g = builder.build_hnsw(emb, M=M, m_levels=m_levels, seed=5, return_stats=True)
    assert g["stats"]["n_overflow"] == 0
    _check_structure(emb, g, M)
    want_v, want_rs = oracle.build_hnsw(emb, g["levels"], M=M, n_levels=2)
    for l in range(2):
        np.testing.assert_array_equal(g["row_splits"][l], want_rs[l])
        np.testing.assert_array_equal(g["values"][l], want_v[l].astype(np.int64))


def test_builder_m32_and_duplicates(oracle):
    """M=32 (the reference's setting: n_cand 96, cap 64) on a corpus with exact duplicate rows (d2 == 0 ties)."""
    from nann_b200 import builder, index as nix
    emb = nix.synthetic_corpus(3000, 128, seed=12)
    emb[100:110] = emb[5]                      # ten copies of one row
    g = builder.build_hnsw(emb, M=32, m_levels=8, seed=6, return_stats=True)
    _check_structure(emb, g, 32)
    want_v, want_rs = oracle.build_hnsw(emb, g["levels"], M=32, n_levels=2)
    for l in range(2):
        np.testing.assert_array_equal(g["row_splits"][l], want_rs[l])
        np.testing.assert_array_equal(g["values"][l], want_v[l].astype(np.int64))


def test_builder_agrees_with_torch_builder():
    from nann_b200 import builder, index as nix
    emb = nix.synthetic_corpus(20000, 128, seed=13)
    t = nix.build_hnsw(emb, M=32, seed=4, device="cuda")
    g = builder.build_hnsw(emb, M=32, seed=4)
    np.testing.assert_array_equal(g["enter_points"], t["enter_points"])
    for l in range(2):
        a = set(zip(np.repeat(np.arange(20000), np.diff(g["row_splits"][l])).tolist(), g["values"][l].tolist()))
        b = set(zip(np.repeat(np.arange(20000), np.diff(t["row_splits"][l])).tolist(), t["values"][l].tolist()))
        assert len(a & b) / max(len(a | b), 1) >= 0.98, (l, len(a), len(b), len(a & b))


def test_builder_files_feed_the_search_path(oracle):
    """end to end: files built on the GPU -> index -> search bit-equal to the oracle on the same files"""
    import nann_b200 as nb
    from nann_b200 import builder, index as nix, scorer_weights as sw
    n, T = 8000, [20, 40, 40, 40, 40, 40]
    emb = nix.synthetic_corpus(n, 128, seed=0)
    ids = nix.synthetic_item_ids(n)
    g = builder.build_hnsw(emb, M=16, m_levels=6, seed=4)
    users = nix.synthetic_queries(emb, 16)
    W = sw.mlp_weights()
    ix = nb.Index.from_arrays(emb, ids, g["enter_points"], g["values"], g["row_splits"])
    got = nb.Searcher(ix, nb.Scorer.mlp(*W), 16, T).search(users, T)
    oix = oracle.Index(emb, ids, g["enter_points"].astype(np.int32), [v.astype(np.int32) for v in g["values"]], g["row_splits"])
    want = oix.search_batch_mlp(oracle.Mlp(*W), users, T, nthreads=0)
    assert np.all(got["status"] == 0)
    np.testing.assert_array_equal(got["ids"], want["ids"])
