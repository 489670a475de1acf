"""The `main.py --job-type test` traversal (NANN_impls/nann/model/model.py:299-362).

CPU: the C oracle against a numpy restatement that uses the same set primitives the reference graph uses
(np.unique / np.setdiff1d are sorted like tf.sets.set_difference; stable descending argsort = tf.math.top_k's
tie rule).  GPU: nann_search_eval_batch against the oracle, bit-exact ids/ranks/scores with the EXACT scorer."""
import numpy as np
import pytest

from tests import util

CASES = [((3, 1, 1), (40, 20, 10), 20),          # clamped ks smaller than the result list
         ((2, 2, 1), (100, 50, 30), 200),        # topk_eval larger than what a level keeps -> n < topk_eval
         ((3, 1, 1), (400, 200, 100), 200),      # the reference's defaults (nann/config.py:52-57)
         ((0, 1, 1), (64, 32, 16), 8)]           # no scoring at level 0: results = level 1's


def numpy_eval(w, score, ns, tk, k):
    nbr = [(w["values"][l], w["row_splits"][l]) for l in range(2)]

    def topk(i, s, kk):
        kk = min(kk, len(i))                                       # tf.reduce_min([k, n])  model.py:268
        o = np.argsort(-s, kind="stable")[:kk]
        return i[o], s[o]

    R = w["ep"].astype(np.int64)
    Rs = score(R)
    tot = len(R)
    R, Rs = topk(R, Rs, tk[2])
    for level in (1, 0):
        visited, C = R.copy(), R.copy()
        for _ in range(ns[level]):
            v, rs = nbr[level]
            flat = np.concatenate([v[rs[c]:rs[c + 1]] for c in C]) if len(C) else np.zeros(0, np.int64)
            nx = np.setdiff1d(np.unique(flat), visited)           # :319-322 (ascending)
            visited = np.union1d(visited, nx)                      # :324
            sx = score(nx)
            tot += len(nx)
            R, Rs = topk(np.concatenate([R, nx]), np.concatenate([Rs, sx]), tk[level])   # :329-331
            C = nx[sx >= Rs[-1]]                                   # :333-334
    return w["item_ids"][R[:k]], Rs[:k], tot


@pytest.mark.parametrize("ns,tk,k", CASES)
def test_oracle_eval_matches_numpy_restatement(oracle, small_world, ns, tk, k):
    w = small_world
    oix = util.oracle_index(oracle, w)
    m = oracle.Mlp(*w["mlp"])
    for q in (0, 7):
        u = w["queries"][q]
        score = lambda ids: m.score(u, w["emb"], np.asarray(ids, np.int32)) if len(ids) else np.zeros(0, np.float32)
        want_ids, want_sc, want_tot = numpy_eval(w, score, ns, tk, k)
        got = oix.search_eval(lambda rnd, ids: score(ids), ns, tk, k)
        assert got["status"] == 0
        assert got["n"] == len(want_ids)
        np.testing.assert_array_equal(got["ids"][:got["n"]], want_ids)
        np.testing.assert_array_equal(got["scores"][:got["n"]].view(np.uint32), want_sc.view(np.uint32))
        assert got["n_scored"] == want_tot
        assert np.all(got["ids"][got["n"]:] == -1)


def test_oracle_eval_rejects_what_the_graph_rejects(oracle, small_world):
    w = small_world
    oix = util.oracle_index(oracle, w)
    zero = lambda rnd, ids: np.zeros(len(ids), np.float32)
    assert oix.search_eval(zero, (3, 1, 2), (40, 20, 10), 20)["status"] != 0     # assert num_scoring[start_level] == 1
    # a single enter point: tf.squeeze -> scalar scores
    one = oracle.Index(w["emb"], w["item_ids"], w["ep"][:1].astype(np.int32),
                       [v.astype(np.int32) for v in w["values"]], w["row_splits"])
    assert one.search_eval(zero, (3, 1, 1), (40, 20, 10), 20)["status"] != 0


@pytest.mark.gpu
@pytest.mark.parametrize("ns,tk,k", CASES)
def test_gpu_eval_search_bit_exact(oracle, small_world, ns, tk, k):
    import nann_b200 as nb
    w = small_world
    oix = util.oracle_index(oracle, w)
    m = oracle.Mlp(*w["mlp"])
    ix = nb.Index.from_arrays(w["emb"], w["item_ids"], w["ep"], w["values"], w["row_splits"])
    sc = nb.Scorer.mlp(*w["mlp"])
    B = 24
    es = nb.EvalSearcher(ix, sc, B, max_top_k_per_level=(400, 200, 100), max_topk_eval=200)
    users = w["queries"][:B]
    before = nb.launch_count()
    got = es.search(users, ns, tk, k)
    assert nb.launch_count() > before
    tot = 0
    for q in range(B):
        u = users[q]
        want = oix.search_eval(lambda rnd, ids: m.score(u, w["emb"], ids) if len(ids) else np.zeros(0, np.float32), ns, tk, k)
        assert got["status"][q] == want["status"] == 0
        assert got["n"][q] == want["n"]
        np.testing.assert_array_equal(got["ids"][q], want["ids"])
        np.testing.assert_array_equal(got["nodes"][q], want["nodes"])
        n = want["n"]
        np.testing.assert_array_equal(got["scores"][q][:n].view(np.uint32), want["scores"][:n].view(np.uint32))
        tot += want["n_scored"]
    assert got["n_scored"] == tot
    # a second call on the same workspace gives the same answer (the candidate bitmap is left clean)
    again = es.search(users, ns, tk, k)
    np.testing.assert_array_equal(again["ids"], got["ids"])
    np.testing.assert_array_equal(again["scores"].view(np.uint32), got["scores"].view(np.uint32))


@pytest.mark.gpu
def test_gpu_eval_search_tensor_precision_and_errors(oracle, small_world):
    import nann_b200 as nb
    w = small_world
    ix = nb.Index.from_arrays(w["emb"], w["item_ids"], w["ep"], w["values"], w["row_splits"])
    sc = nb.Scorer.mlp(*w["mlp"])
    es = nb.EvalSearcher(ix, sc, 16)
    users = w["queries"][:16]
    exact = es.search(users)
    sc.set_precision(nb.SCORER_TENSOR)
    tens = es.search(users)
    assert np.all(tens["status"] == 0) and np.array_equal(tens["n"], exact["n"])
    # same items up to rank flips between scores closer than the tensor path's tolerance
    for q in range(16):
        n = exact["n"][q]
        common = len(set(exact["ids"][q][:n].tolist()) & set(tens["ids"][q][:n].tolist()))
        assert common >= n - 2
        pa = dict(zip(exact["ids"][q][:n].tolist(), exact["scores"][q][:n].tolist()))
        for i, s in zip(tens["ids"][q][:n].tolist(), tens["scores"][q][:n].tolist()):
            if i in pa:
                assert abs(pa[i] - s) <= 1e-5
    with pytest.raises(nb.NannError):
        es.search(users, (3, 1, 2))                       # assert num_scoring[start_level] == 1 (model.py:347)
    with pytest.raises(nb.NannError):
        es.search(users, top_k_per_level=(401, 200, 100))  # beyond the workspace
