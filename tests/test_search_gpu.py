"""Traversal parity (the exec.pb dataflow, build_opt_graph.py:109-149) on the GPU.

EXACT scorer: the whole search is bit-exact against the oracle -- item ids, ranks, scores and the
per-round candidate lists.  Independently of the scorer, the integer traversal is checked to be
bit-exact GIVEN the GPU's own score arrays (oracle re-run fed with the traced GPU scores)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nb():
    import nann_b200
    return nann_b200


@pytest.fixture(scope="module")
def world(nb, small_world, oracle):
    from tests import util
    w = dict(small_world)
    w["ix"] = nb.Index.from_arrays(w["emb"], w["item_ids"], w["ep"], w["values"], w["row_splits"])
    w["scorer"] = nb.Scorer.mlp(*w["mlp"])
    w["oix"] = util.oracle_index(oracle, w)
    w["omlp"] = oracle.Mlp(*w["mlp"])
    return w


def _oracle_batch(world, users, T):
    return world["oix"].search_batch_mlp(world["omlp"], users, T, nthreads=0)


@pytest.mark.parametrize("B", [1, 3, 64])
def test_search_bit_exact_vs_oracle(nb, world, B):
    T = world["T"]
    s = nb.Searcher(world["ix"], world["scorer"], 64, T)
    users = world["queries"][:B]
    got = s.search(users, T)
    want = _oracle_batch(world, users, T)
    assert np.all(got["status"] == 0) and np.all(want["status"] == 0)
    np.testing.assert_array_equal(got["ids"], want["ids"])
    np.testing.assert_array_equal(got["scores"].view(np.uint32), want["scores"].view(np.uint32))
    assert got["n_scored"].sum() == want["n_scored"]


def test_search_trace_matches_oracle_rounds(nb, oracle, world):
    T = world["T"]
    s = nb.Searcher(world["ix"], world["scorer"], 8, T)
    s.set_trace(True)
    users = world["queries"][8:16]
    got = s.search(users, T)
    for q in range(8):
        u = users[q]
        ref = world["oix"].search(lambda r, ids: world["omlp"].score(u, world["emb"], ids), T, trace=True)
        for r in range(5):
            ids, sc = s.trace(q, r)
            np.testing.assert_array_equal(ids, ref["trace"][r][0], err_msg=f"q{q} round{r}")
            np.testing.assert_array_equal(sc.view(np.uint32), ref["trace"][r][1].view(np.uint32))
        np.testing.assert_array_equal(got["ids"][q], ref["ids"])
        np.testing.assert_array_equal(s.nodes(8, T[5])[q], ref["nodes"])


def test_traversal_bit_exact_given_gpu_scores(nb, world):
    """Scorer-independent check (used for the tensor-core path too): feed the oracle traversal
    with the GPU's traced scores; ids/ranks must match exactly."""
    T = world["T"]
    s = nb.Searcher(world["ix"], world["scorer"], 4, T)
    s.set_trace(True)
    users = world["queries"][20:24]
    got = s.search(users, T)
    for q in range(4):
        traced = [s.trace(q, r) for r in range(5)]

        def score(r, ids, traced=traced):
            tid, tsc = traced[r]
            np.testing.assert_array_equal(ids, tid)
            return tsc

        ref = world["oix"].search(score, T)
        np.testing.assert_array_equal(got["ids"][q], ref["ids"])
        np.testing.assert_array_equal(got["scores"][q].view(np.uint32), ref["scores"].view(np.uint32))


def test_opwise_path_equals_fused(nb, world):
    """One C-ABI call per exec.pb node (what the TF shim runs) == the fused batched call."""
    import torch
    T = world["T"]
    s = nb.Searcher(world["ix"], world["scorer"], 2, T)
    users = world["queries"][30:32]
    fused = s.search(users, T)
    emb_dev = torch.from_numpy(world["emb"]).cuda()          # HugeConst's cached device copy
    vals = [v.astype(np.int32) for v in world["values"]]      # build_opt_graph.py:87 narrows to int32
    for q in range(2):
        ids, sc = nb.retrieve_opwise(world["scorer"], users[q], T, emb_dev, world["item_ids"], world["ep"],
                                     vals, world["row_splits"])
        assert ids.shape == (1, T[5])
        np.testing.assert_array_equal(ids[0], fused["ids"][q])
        np.testing.assert_array_equal(sc.view(np.uint32), fused["scores"][q].view(np.uint32))


def test_level_topn_is_a_runtime_input(nb, world):
    """beam widths change per call without rebuilding anything (build_opt_graph.py:75)."""
    s = nb.Searcher(world["ix"], world["scorer"], 16, [60, 120, 120, 120, 120, 120])
    users = world["queries"][:16]
    for T in ([50, 100, 100, 100, 100, 100], [60, 120, 80, 40, 20, 120], [10, 10, 10, 10, 10, 5], [50, 100, 100, 100, 100, 0]):
        got = s.search(users, T)
        want = _oracle_batch(world, users, T)
        np.testing.assert_array_equal(got["status"], want["status"])
        np.testing.assert_array_equal(got["ids"], want["ids"])
    with pytest.raises(nb.NannError):
        s.search(users, [61, 10, 10, 10, 10, 10])            # above the workspace maximum
    with pytest.raises(nb.NannError):
        s.search(users, [10, 10, 10, -1, 10, 10])            # "Need k >= 0"


def test_fewer_than_k_candidates_fails_that_query_only(nb, oracle, world):
    n_ep = len(world["ep"])
    T = [n_ep + 1, 10, 10, 10, 10, 10]                       # TopKV2: input must have at least k columns
    s = nb.Searcher(world["ix"], world["scorer"], 4, T)
    got = s.search(world["queries"][:4], T)
    want = _oracle_batch(world, world["queries"][:4], T)
    assert np.all(got["status"] == nb._lib.INVALID_ARGUMENT) and np.all(want["status"] == oracle.INVALID_ARGUMENT)
    assert np.all(got["ids"] == -1) and got["n_failed"] == 4
    T2 = [20, 30, 400, 100, 100, 50]                         # fails later, at the first level-0 top-k, for SOME queries
    s2 = nb.Searcher(world["ix"], world["scorer"], 64, T2)
    got = s2.search(world["queries"], T2)
    want = _oracle_batch(world, world["queries"], T2)
    np.testing.assert_array_equal(got["status"], want["status"])
    ok = want["status"] == 0
    np.testing.assert_array_equal(got["ids"][ok], want["ids"][ok])


def test_determinism_and_batch_composition(nb, world):
    T = world["T"]
    s = nb.Searcher(world["ix"], world["scorer"], 64, T)
    a = s.search(world["queries"], T)
    b = s.search(world["queries"], T)
    np.testing.assert_array_equal(a["ids"], b["ids"])
    perm = np.random.default_rng(0).permutation(64)
    c = s.search(world["queries"][perm], T)
    np.testing.assert_array_equal(c["ids"], a["ids"][perm])   # a query's result does not depend on its batch mates
    import torch
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        d = s.search(torch.from_numpy(world["queries"]).cuda(), T, stream=st)   # device-resident users, side stream
    np.testing.assert_array_equal(d["ids"], a["ids"])


def test_index_load_from_files_and_dtype_widths(nb, world):
    """Appendix-C files: i64 values as written by build_hnsw_index.py, or i32 after exec.pb's in-place cast."""
    T = world["T"]
    ix2 = nb.Index.load(world["embs_dir"], world["index_dir"])
    assert ix2.n_items == world["emb"].shape[0] and ix2.dim == 128 and ix2.n_enter_points == len(world["ep"])
    a = nb.Searcher(ix2, world["scorer"], 8, T).search(world["queries"][:8], T)
    ix3 = nb.Index.from_arrays(world["emb"], world["item_ids"], world["ep"].astype(np.int32),
                               [v.astype(np.int32) for v in world["values"]], world["row_splits"])
    b = nb.Searcher(ix3, world["scorer"], 8, T).search(world["queries"][:8], T)
    c = nb.Searcher(world["ix"], world["scorer"], 8, T).search(world["queries"][:8], T)
    np.testing.assert_array_equal(a["ids"], c["ids"])
    np.testing.assert_array_equal(b["ids"], c["ids"])


def test_index_validation(nb, world):
    bad_vals = [world["values"][0].copy(), world["values"][1]]
    bad_vals[0][5] = world["emb"].shape[0]                                     # neighbour id out of range
    with pytest.raises(nb.NannError):
        nb.Index.from_arrays(world["emb"], world["item_ids"], world["ep"], bad_vals, world["row_splits"])
    with pytest.raises(nb.NannError):
        nb.Index.from_arrays(world["emb"], world["item_ids"], world["ep"][::-1].copy(), world["values"], world["row_splits"])
    rs = [world["row_splits"][0].copy(), world["row_splits"][1]]
    rs[0][-1] += 1
    with pytest.raises(nb.NannError):
        nb.Index.from_arrays(world["emb"], world["item_ids"], world["ep"], world["values"], rs)


def test_attention_scorer_search_config1_shapes(nb, oracle):
    """config 1 plumbing: d=64 f16 table, user [50,64], attention scorer; traversal bit-exact given
    the GPU scores, scores within 1e-5 of the oracle scorer."""
    from nann_b200 import index as nix, scorer_weights as sw
    from tests import util
    w = util.build_world(tag="c1", n=3000, d=64, M=16, m_levels=6, seed=4, n_cand=40, device="cpu")
    emb16 = w["emb"].astype(np.float16)
    blob = sw.attention_blob(seed=3)
    sc = nb.Scorer.attention(blob)
    ix = nb.Index.from_arrays(emb16, w["item_ids"], w["ep"], w["values"], w["row_splits"])     # f16 table, widened once
    T = [20, 40, 40, 40, 40, 40]
    s = nb.Searcher(ix, sc, 4, T)
    s.set_trace(True)
    rng = np.random.default_rng(0)
    users = (0.01 * rng.random((4, 3200))).astype(np.float16).astype(np.float32)              # comm_seq f16[1,3200]
    got = s.search(users, T)
    assert np.all(got["status"] == 0)
    emb32 = emb16.astype(np.float32)
    oa = oracle.Attn(blob)
    oix = oracle.Index(emb32, w["item_ids"], w["ep"].astype(np.int32), [v.astype(np.int32) for v in w["values"]], w["row_splits"])
    for q in range(4):
        traced = [s.trace(q, r) for r in range(5)]
        for r in range(5):
            want = oa.score(users[q], emb32, traced[r][0])
            assert np.abs(want - traced[r][1]).max() <= 1e-5
        ref = oix.search(lambda r, ids, t=traced: t[r][1], T)
        np.testing.assert_array_equal(got["ids"][q], ref["ids"])


def test_executor_closed_loop(nb, world):
    """blaze-benchmark role: predictor_num searchers/streams, dynamic batching, cppmetrics-style report."""
    from nann_b200 import harness
    T = world["T"]
    r1 = harness.run_benchmark(world["ix"], world["scorer"], T, world["queries"], predictor_num=2, bench_thread_count=2,
                               duration=1.0, max_batch_size=1)
    assert r1["failures"] == 0 and r1["throughput_count"] > 0
    assert r1["batchsize"]["max"] == 1 and r1["latency_us"]["median"] > 0          # the reference's batch=1 mode
    r2 = harness.run_benchmark(world["ix"], world["scorer"], T, world["queries"], predictor_num=2, bench_thread_count=2,
                               duration=1.0, max_batch_size=32)
    assert r2["failures"] == 0 and r2["batchsize"]["max"] <= 32 and r2["batchsize"]["mean"] > 1
    assert r2["throughput"] > r1["throughput"]                                      # batching is the throughput lever
    assert r2["latency_us"]["p99"] >= r2["latency_us"]["median"] >= r2["latency_us"]["min"]
    r3 = harness.run_benchmark(world["ix"], world["scorer"], T, world["queries"], predictor_num=1, bench_thread_count=1,
                               duration=1.0, qps=200, max_batch_size=8)
    assert 100 <= r3["throughput"] <= 260                                           # paced load is respected


def test_tensor_core_scorer_search_parity(nb, oracle, world):
    """TENSOR scorer: (1) every traced score within 1e-5 of the oracle's exact score, (2) the integer
    traversal bit-exact GIVEN those scores, (3) the final top-k agrees with the exact path except
    where two scores are closer than the tolerance."""
    T = world["T"]
    sc = nb.Scorer.mlp(*world["mlp"])
    sc.set_precision(nb.SCORER_TENSOR)
    s = nb.Searcher(world["ix"], sc, 16, T)
    s.set_trace(True)
    users = world["queries"][32:48]
    got = s.search(users, T)
    assert np.all(got["status"] == 0)
    exact = _oracle_batch(world, users, T)
    overlap = 0
    for q in range(16):
        traced = [s.trace(q, r) for r in range(5)]
        for r in range(5):
            want = world["omlp"].score(users[q], world["emb"], traced[r][0])
            assert np.abs(want - traced[r][1]).max() <= 1e-5
        ref = world["oix"].search(lambda r, ids, t=traced: t[r][1], T)
        np.testing.assert_array_equal(got["ids"][q], ref["ids"])
        np.testing.assert_array_equal(got["scores"][q].view(np.uint32), ref["scores"].view(np.uint32))
        overlap += len(set(got["ids"][q].tolist()) & set(exact["ids"][q].tolist()))
    assert overlap >= 0.99 * 16 * T[5]


def _adversarial_index(n=3000, seed=11):
    """A graph built to stress the visited filter: heavy duplication inside and across neighbour lists (every list
    draws from a small hub set, with repeats), empty rows, and lists that are entirely visited after one round."""
    rng = np.random.default_rng(seed)
    hubs = rng.choice(n, 90, replace=False)
    vals, rs = [[], []], [[0], [0]]
    for l, deg in ((0, 64), (1, 32)):
        for i in range(n):
            k = 0 if i % 17 == 0 else int(rng.integers(1, deg + 1))
            row = np.where(rng.random(k) < 0.8, rng.choice(hubs, k), rng.integers(0, n, k))
            if k > 3:
                row[1] = row[0]                      # duplicates inside one list
            vals[l].extend(row.tolist())
            rs[l].append(len(vals[l]))
    emb = (rng.standard_normal((n, 128)) / 11.3).astype(np.float32)
    return dict(emb=emb, item_ids=rng.permutation(n).astype(np.int64), ep=np.sort(rng.choice(n, 120, replace=False)).astype(np.int64),
                values=[np.asarray(v, np.int64) for v in vals], row_splits=[np.asarray(r, np.int64) for r in rs])


def test_expand_filter_cta_kernel_on_adversarial_graph(nb, oracle, world):
    """The CTA-per-query expand+filter (shared-memory min-position hash) against the oracle's serial loop on a graph
    with heavy id duplication; bit-exact ids, ranks and scores, and identical per-round candidate counts."""
    from nann_b200 import scorer_weights as sw
    w = _adversarial_index()
    W = sw.mlp_weights(seed=3)
    ix = nb.Index.from_arrays(w["emb"], w["item_ids"], w["ep"], w["values"], w["row_splits"])
    sc = nb.Scorer.mlp(*W)
    T = [40, 60, 60, 60, 60, 100]
    rng = np.random.default_rng(5)
    users = (rng.standard_normal((9, 128)) / 11.3).astype(np.float32)
    got = nb.Searcher(ix, sc, 9, T).search(users, T)
    oix = oracle.Index(w["emb"], w["item_ids"], w["ep"].astype(np.int32), [v.astype(np.int32) for v in w["values"]], w["row_splits"])
    want = oix.search_batch_mlp(oracle.Mlp(*W), users, T, nthreads=0)
    np.testing.assert_array_equal(got["status"], want["status"])
    ok = want["status"] == 0
    assert ok.sum() >= 5
    np.testing.assert_array_equal(got["ids"][ok], want["ids"][ok])
    np.testing.assert_array_equal(got["scores"][ok].view(np.uint32), want["scores"][ok].view(np.uint32))
    assert got["n_scored"].sum() == want["n_scored"]


def test_mid_size_corpus_large_ragged_batch(nb, oracle):
    """40 k items (M=32 like the bench), ef=400-shaped beam widths, a batch of 300 (not a power of two, several waves of
    CTAs per kernel): every query of the batch is checked bit-exact against the oracle in EXACT mode; the tensor-core
    path returns the same items up to near-tie flips."""
    from tests import util
    from nann_b200 import scorer_weights as sw
    w = util.build_world(40000, M=32, m_levels=32, seed=4, device="cuda", tag="mid")
    W = sw.mlp_weights(seed=3)
    T = [30, 120, 240, 240, 240, 200]
    from nann_b200 import index as nix
    users = nix.synthetic_queries(w["emb"], 300, seed=7)
    ix = nb.Index.from_arrays(w["emb"], w["item_ids"], w["ep"], w["values"], w["row_splits"])
    sc = nb.Scorer.mlp(*W)
    se = nb.Searcher(ix, sc, 300, T)
    got = se.search(users, T)
    import os
    want = util.oracle_index(oracle, w).search_batch_mlp(oracle.Mlp(*W), users, T, nthreads=min(16, os.cpu_count() or 1))
    np.testing.assert_array_equal(got["status"], want["status"])
    assert np.all(want["status"] == 0)
    np.testing.assert_array_equal(got["ids"], want["ids"])
    np.testing.assert_array_equal(got["scores"].view(np.uint32), want["scores"].view(np.uint32))
    assert got["n_scored"].sum() == want["n_scored"]
    sc.set_precision(nb.SCORER_TENSOR)
    tens = se.search(users, T)
    assert np.all(tens["status"] == 0)
    overlap = np.mean([len(set(a.tolist()) & set(b.tolist())) / T[5] for a, b in zip(tens["ids"], want["ids"])])
    assert overlap >= 0.995


def test_two_shards_on_one_device_merge_like_the_oracle(nb, oracle, world):
    """Row-sharded search in one process: two shards of the small corpus with their own HNSW, per-shard
    beams, nann_merge_topk; the merged result equals the stable merge (score desc, ties -> lower shard, then rank)
    of the per-shard ORACLE results."""
    from nann_b200 import index as nix
    from nann_b200.distributed import shard_bounds, shard_level_topn
    T = world["T"]
    Ts = shard_level_topn(T, 2)
    users = world["queries"][:8]
    sc_l, id_l, osc_l, oid_l = [], [], [], []
    for r in range(2):
        lo, hi = shard_bounds(world["emb"].shape[0], 2, r)
        emb, ids = world["emb"][lo:hi], world["item_ids"][lo:hi]
        g = nix.build_hnsw(emb, M=16, m_levels=6, seed=4 + r, n_cand=40, device="cpu")
        ix = nb.Index.from_arrays(emb, ids, g["enter_points"], g["values"], g["row_splits"])
        res = nb.Searcher(ix, world["scorer"], 8, Ts).search(users, Ts)
        assert np.all(res["status"] == 0)
        sc_l.append(res["scores"]); id_l.append(res["ids"])
        oix = oracle.Index(emb, ids, g["enter_points"].astype(np.int32), [v.astype(np.int32) for v in g["values"]], g["row_splits"])
        o = oix.search_batch_mlp(world["omlp"], users, Ts, nthreads=0)
        osc_l.append(o["scores"]); oid_l.append(o["ids"])
    m_sc, m_id = nb.merge_topk(np.stack(sc_l), np.stack(id_l), T[5])
    for q in range(8):
        cs = np.concatenate([osc_l[0][q], osc_l[1][q]])
        ci = np.concatenate([oid_l[0][q], oid_l[1][q]])
        o = np.argsort(-cs, kind="stable")[:T[5]]
        np.testing.assert_array_equal(m_id[q], ci[o])
        np.testing.assert_array_equal(m_sc[q].view(np.uint32), cs[o].view(np.uint32))


def test_shard_group_exchange_in_library(nb, oracle, world):
    """nann_search_sharded: two shard members in ONE process on one device (peer windows = plain device pointers).
    The final top-k kernel pushes its records into both windows, the merge kernel runs on the group's stream; six
    sequences exercise the double-buffered windows and the done-flag back-pressure.  The merged result equals the
    stable merge (score desc, ties -> lower shard, then lower rank) of the per-shard ORACLE results, bit for bit."""
    import torch
    from nann_b200 import index as nix
    from nann_b200.distributed import ShardGroup, shard_bounds, shard_level_topn
    T = world["T"]
    Ts = shard_level_topn(T, 2)
    B, n_seq = 8, 6
    shards, want = [], []
    for r in range(2):
        lo, hi = shard_bounds(world["emb"].shape[0], 2, r)
        emb, ids = world["emb"][lo:hi], world["item_ids"][lo:hi]
        g = nix.build_hnsw(emb, M=16, m_levels=6, seed=4 + r, n_cand=40, device="cpu")
        ix = nb.Index.from_arrays(emb, ids, g["enter_points"], g["values"], g["row_splits"])
        shards.append((ix, nb.Searcher(ix, world["scorer"], B, Ts)))
        oix = oracle.Index(emb, ids, g["enter_points"].astype(np.int32), [v.astype(np.int32) for v in g["values"]], g["row_splits"])
        want.append([oix.search_batch_mlp(world["omlp"], world["queries"][i * B:(i + 1) * B], Ts, nthreads=0) for i in range(n_seq)])
    members = [ShardGroup(r, 2, B, Ts[5]) for r in range(2)]
    ShardGroup.connect_local(members)
    outs = [[(torch.empty((B, T[5]), dtype=torch.int64, device="cuda"), torch.empty((B, T[5]), dtype=torch.float32, device="cuda"),
              torch.empty((B,), dtype=torch.int32, device="cuda")) for _ in range(n_seq)] for _ in range(2)]
    before = nb.launch_count()
    all_users = torch.from_numpy(world["queries"][:n_seq * B]).cuda()
    side = torch.cuda.Stream()                  # not the legacy default stream: that one serialises with everything
    torch.cuda.synchronize()
    for i in range(n_seq):
        for r in range(2):                      # one host thread drives both members: push both, then merge both
            members[r].push(shards[r][1], all_users[i * B:(i + 1) * B], Ts, stream=side)
        for r in range(2):
            members[r].merge(T[5], *outs[r][i])
    for m in members:
        m.wait()
    torch.cuda.synchronize()
    assert nb.launch_count() - before >= n_seq * 2 * 20
    for i in range(n_seq):
        for q in range(B):
            cs = np.concatenate([want[0][i]["scores"][q], want[1][i]["scores"][q]])
            ci = np.concatenate([want[0][i]["ids"][q], want[1][i]["ids"][q]])
            o = np.argsort(-cs, kind="stable")[:T[5]]
            for r in range(2):                  # every rank ends with the same global list
                np.testing.assert_array_equal(outs[r][i][0][q].cpu().numpy(), ci[o])
                np.testing.assert_array_equal(outs[r][i][1][q].cpu().numpy().view(np.uint32), cs[o].view(np.uint32))
        assert all(int(outs[r][i][2].sum()) == 0 for r in range(2))


def test_shard_group_failed_query_fails_everywhere(nb, oracle, world):
    """A query whose search fails on ONE shard (TopKV2 n < k there) fails on every rank alike: combined status,
    ids -1 -- never a half-merged list."""
    import torch
    from nann_b200 import index as nix
    from nann_b200.distributed import ShardGroup, shard_bounds
    B = 4
    Ts = [[20, 40, 40, 40, 40, 40], [20, 40, 40, 40, 40, 40]]
    users = torch.from_numpy(world["queries"][:B]).cuda()
    members = [ShardGroup(r, 2, B, 40) for r in range(2)]
    ShardGroup.connect_local(members)
    searchers, keep = [], []
    for r in range(2):
        lo, hi = shard_bounds(world["emb"].shape[0], 2, r)
        if r == 1:
            hi = lo + 300                        # a tiny shard: level_topn[0]=20 > its enter points -> InvalidArgument
        emb, ids = world["emb"][lo:hi], world["item_ids"][lo:hi]
        g = nix.build_hnsw(emb, M=16, m_levels=6, seed=4 + r, n_cand=40, device="cpu")
        if r == 1:
            assert len(g["enter_points"]) < 20
        ix = nb.Index.from_arrays(emb, ids, g["enter_points"], g["values"], g["row_splits"])
        keep.append(ix)
        searchers.append(nb.Searcher(ix, world["scorer"], B, Ts[r]))
    outs = [(torch.zeros((B, 40), dtype=torch.int64, device="cuda"), torch.zeros((B, 40), dtype=torch.float32, device="cuda"),
             torch.zeros((B,), dtype=torch.int32, device="cuda")) for _ in range(2)]
    side = torch.cuda.Stream()
    torch.cuda.synchronize()
    for r in range(2):
        members[r].push(searchers[r], users, Ts[r], stream=side)
    for r in range(2):
        members[r].merge(40, *outs[r])
    for m in members:
        m.wait()
    torch.cuda.synchronize()
    for r in range(2):
        assert np.all(outs[r][2].cpu().numpy() == 3)
        assert np.all(outs[r][0].cpu().numpy() == -1)


def test_small_batches_replay_a_cuda_graph_bit_identically(nb, world):
    """batch <= 32 on a capturable stream: call 1 runs eagerly, call 2 captures the launch sequence into a CUDA graph,
    later calls replay it.  Every call must return the oracle's ids and scores bit for bit, for different queries,
    batch sizes and level_topn, in EXACT and TENSOR precision (the latter against its own eager result)."""
    import torch
    T = world["T"]
    T2 = [max(t // 2, 8) for t in T]
    se = nb.Searcher(world["ix"], world["scorer"], 8, T)
    side = torch.cuda.Stream()
    launches = []
    for rep in range(5):
        for B, Tq in ((1, T), (3, T), (1, T2)):
            users = world["queries"][rep * 8:rep * 8 + B]
            l0 = nb.launch_count()
            got = se.search(users, Tq, stream=side)
            launches.append(nb.launch_count() - l0)
            want = _oracle_batch(world, users, Tq)
            np.testing.assert_array_equal(got["ids"], want["ids"])
            np.testing.assert_array_equal(got["scores"].view(np.uint32), want["scores"].view(np.uint32))
            assert got["n_scored"].sum() == want["n_scored"]
    assert len(set(launches)) == 1                      # replays are counted kernel by kernel, like eager calls
    # tensor-core scorer: replayed == eager, bit for bit
    world["scorer"].set_precision(nb.SCORER_TENSOR)
    try:
        users = world["queries"][:2]
        runs = [se.search(users, T, stream=side) for _ in range(4)]
        for r in runs[1:]:
            np.testing.assert_array_equal(r["ids"], runs[0]["ids"])
            np.testing.assert_array_equal(r["scores"].view(np.uint32), runs[0]["scores"].view(np.uint32))
    finally:
        world["scorer"].set_precision(nb.SCORER_EXACT)


def test_host_calls_without_a_stream_replay_graphs_too(nb, world):
    """The reference's operating mode: batch 1, host arrays in and out, no stream argument (the NULL stream cannot be
    captured).  Such calls run on a stream of the searcher's own and replay the captured sequence; every call returns the
    oracle's result bit for bit, mixed with calls on a caller stream and with device inputs on the NULL stream."""
    import time
    import torch
    T = world["T"]
    se = nb.Searcher(world["ix"], world["scorer"], 8, T)
    side = torch.cuda.Stream()
    for rep in range(6):
        for B in (1, 2):
            users = world["queries"][40 + rep * 2:40 + rep * 2 + B]
            want = _oracle_batch(world, users, T)
            for got in (se.search(users, T), se.search(users, T, stream=side), se.search(torch.from_numpy(users).cuda(), T)):
                np.testing.assert_array_equal(got["ids"], want["ids"])
                np.testing.assert_array_equal(got["scores"].view(np.uint32), want["scores"].view(np.uint32))
    users = world["queries"][:1]

    def median_ms(**kw):
        ts = []
        for _ in range(30):
            t0 = time.perf_counter()
            se.search(users, T, **kw)
            ts.append(time.perf_counter() - t0)
        return 1e3 * sorted(ts)[len(ts) // 2]

    # informational only (no assertion on time): on this small corpus in EXACT precision a batch-1 call is ~1 ms of fp32
    # scoring, so the replay saves little; the launch-bound case is the tensor path (scripts/sweep.py: 0.42 ms at batch 1)
    replay = median_ms()
    se.set_trace(True)                                   # tracing disables the graph path: the eager launch sequence
    eager = median_ms()
    se.set_trace(False)
    print(f"batch-1 host call without a stream: {replay:.3f} ms replayed, {eager:.3f} ms eager (with trace copies)")


@pytest.mark.parametrize("precision", ["exact", "tensor"])
def test_distributed_scoring_world1_is_bit_identical_to_the_plain_search(nb, world, precision):
    """nann_search_distributed with a group of ONE: every candidate is owned by this rank, but the whole exchange path
    runs (bucket -> own request window -> scorer on the pseudo-queries -> return -> unbucket, flags and one-warp waits).
    Ids and scores must equal nann_search_batch bit for bit, over several calls and batch sizes."""
    import torch
    from nann_b200.distributed import DistGroup
    T = world["T"]
    n = world["emb"].shape[0]
    sc = nb.Scorer.mlp(*world["mlp"])
    if precision == "tensor":
        sc.set_precision(nb.SCORER_TENSOR)
    ref = nb.Searcher(world["ix"], sc, 16, T)
    ix = nb.Index.from_arrays_sharded(n, world["emb"], 0, world["item_ids"], world["ep"], world["values"], world["row_splits"])
    se = nb.Searcher(ix, sc, 16, T)
    grp = DistGroup(se, 0, 1)
    side = torch.cuda.Stream()
    for B in (16, 5, 16):
        users = world["queries"][:B]
        sc_h, id_h, st_h = grp.search(users, T)                       # host outputs
        want = ref.search(users, T)
        np.testing.assert_array_equal(id_h, want["ids"])
        np.testing.assert_array_equal(sc_h.view(np.uint32), want["scores"].view(np.uint32))
        assert np.all(st_h == 0)
        o = (torch.empty((B, T[5]), dtype=torch.int64, device="cuda"), torch.empty((B, T[5]), dtype=torch.float32, device="cuda"),
             torch.empty((B,), dtype=torch.int32, device="cuda"))
        grp.search(torch.from_numpy(users).cuda(), T, *o, stream=side)   # device outputs, enqueue only
        torch.cuda.synchronize()
        grp.check()
        np.testing.assert_array_equal(o[0].cpu().numpy(), want["ids"])
        np.testing.assert_array_equal(o[1].cpu().numpy().view(np.uint32), want["scores"].view(np.uint32))


@pytest.mark.parametrize("precision", ["exact", "tensor"])
def test_distributed_scoring_two_members_two_threads(nb, world, precision):
    """Two members on ONE device, each driven by its own host thread and stream (the arrangement of a server with one
    thread per GPU): the graph is replicated, the embedding table row-sharded, each member traverses ITS queries and scores
    the candidates it owns for both.  Ids and scores equal nann_search_batch on the unsharded index bit for bit."""
    import threading
    import torch
    from nann_b200.distributed import DistGroup
    T = world["T"]
    n = world["emb"].shape[0]
    B, n_seq, G = 6, 3, 2
    sc = nb.Scorer.mlp(*world["mlp"])
    if precision == "tensor":
        sc.set_precision(nb.SCORER_TENSOR)
    ref = nb.Searcher(world["ix"], sc, G * B, T)
    per = -(-n // G)
    members, keep = [], []
    for r in range(G):
        lo, hi = r * per, min((r + 1) * per, n)
        ix = nb.Index.from_arrays_sharded(n, world["emb"][lo:hi], lo, world["item_ids"], world["ep"], world["values"], world["row_splits"])
        se = nb.Searcher(ix, sc, B, T)
        with pytest.raises(nb.NannError):              # a slice of the table cannot be searched on its own
            se.search(world["queries"][:B], T)
        keep.append((ix, se))
        members.append(DistGroup(se, r, G))
    DistGroup.connect_local(members)
    streams = [torch.cuda.Stream() for _ in range(G)]
    res = [[None] * n_seq for _ in range(G)]
    errs = []

    def drive(r):
        try:
            for i in range(n_seq):                     # rank r owns queries [r*B, (r+1)*B) of the global batch i
                res[r][i] = members[r].search(world["queries"][(i * G + r) * B:(i * G + r + 1) * B], T, stream=streams[r])
        except Exception as e:                         # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=drive, args=(r,)) for r in range(G)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for i in range(n_seq):
        want = ref.search(world["queries"][i * G * B:(i + 1) * G * B], T)
        assert np.all(want["status"] == 0)
        for r in range(G):
            sc_h, id_h, st_h = res[r][i]
            np.testing.assert_array_equal(id_h, want["ids"][r * B:(r + 1) * B])
            np.testing.assert_array_equal(sc_h.view(np.uint32), want["scores"][r * B:(r + 1) * B].view(np.uint32))
            assert np.all(st_h == 0)


def test_distributed_scoring_attention_scorer(nb):
    """The reference's own scorer (config 1: d=64 f16 table, user [50,64]) through distributed scoring: the key projections
    AND the raw user sequences travel to the owners.  A group of one, then two members on two threads; ids and scores equal
    nann_search_batch on the unsharded index bit for bit."""
    import threading
    import torch
    from nann_b200 import scorer_weights as sw
    from nann_b200.distributed import DistGroup
    from tests import util
    w = util.build_world(tag="c1", n=3000, d=64, M=16, m_levels=6, seed=4, n_cand=40, device="cpu")
    emb16 = w["emb"].astype(np.float16)
    n = emb16.shape[0]
    sc = nb.Scorer.attention(sw.attention_blob(seed=3))
    T = [20, 40, 40, 40, 40, 40]
    B, G = 3, 2
    rng = np.random.default_rng(0)
    users = (0.01 * rng.random((2 * G * B, 3200))).astype(np.float16).astype(np.float32)
    ix_full = nb.Index.from_arrays(emb16, w["item_ids"], w["ep"], w["values"], w["row_splits"])
    ref = nb.Searcher(ix_full, sc, G * B, T)
    # a group of one
    ix1 = nb.Index.from_arrays_sharded(n, emb16, 0, w["item_ids"], w["ep"], w["values"], w["row_splits"])
    se1 = nb.Searcher(ix1, sc, G * B, T)
    g1 = DistGroup(se1, 0, 1)
    want = ref.search(users[:G * B], T)
    assert np.all(want["status"] == 0)
    sc_h, id_h, st_h = g1.search(users[:G * B], T)
    np.testing.assert_array_equal(id_h, want["ids"])
    np.testing.assert_array_equal(sc_h.view(np.uint32), want["scores"].view(np.uint32))
    # two members, two threads
    per = -(-n // G)
    members, keep = [], []
    for r in range(G):
        lo, hi = r * per, min((r + 1) * per, n)
        ix = nb.Index.from_arrays_sharded(n, emb16[lo:hi], lo, w["item_ids"], w["ep"], w["values"], w["row_splits"])
        se = nb.Searcher(ix, sc, B, T)
        keep.append((ix, se))
        members.append(DistGroup(se, r, G))
    DistGroup.connect_local(members)
    streams = [torch.cuda.Stream() for _ in range(G)]
    res = [[None] * 2 for _ in range(G)]
    errs = []

    def drive(r):
        try:
            for i in range(2):
                res[r][i] = members[r].search(users[(i * G + r) * B:(i * G + r + 1) * B], T, stream=streams[r])
        except Exception as e:                         # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=drive, args=(r,)) for r in range(G)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for i in range(2):
        want = ref.search(users[i * G * B:(i + 1) * G * B], T)
        for r in range(G):
            sc_h, id_h, st_h = res[r][i]
            np.testing.assert_array_equal(id_h, want["ids"][r * B:(r + 1) * B])
            np.testing.assert_array_equal(sc_h.view(np.uint32), want["scores"][r * B:(r + 1) * B].view(np.uint32))
            assert np.all(st_h == 0)
