"""Pins the CPU oracle against the reference's own vectors (tests/golden/kat_ops.json) and against
independent numpy restatements.  CPU only."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat_ops.json")))


def test_group_gather_golden(oracle):
    for c in KAT["group_gather"]:
        dt = np.dtype(c["dtype"])
        v, rs = oracle.group_gather(np.array(c["params_values"], dt), c["params_row_splits"],
                                    c["indices_values"], c["indices_row_splits"], unique=c["unique"])
        assert rs.tolist() == c["ret_row_splits"], c["source"]
        if "ret_values" in c:
            assert v.tolist() == c["ret_values"], c["source"]
        else:
            for g, want in enumerate(c["ret_sets"]):
                got = v[rs[g]:rs[g + 1]].tolist()
                assert sorted(got) == want and len(set(got)) == len(got), c["source"]


def test_group_gather_i32_matches_i64(oracle):
    c = KAT["group_gather"][0]
    a = oracle.group_gather(np.array(c["params_values"], np.int32), c["params_row_splits"],
                            c["indices_values"], c["indices_row_splits"])
    assert a[0].dtype == np.int32 and a[0].tolist() == c["ret_values"]


@pytest.mark.parametrize("rs,code", [([], 1), ([1, 3], 2), ([0, 2], 3)])
def test_ragged_validation_codes(oracle, rs, code):
    # GroupGather_kernel.cc:9-16
    assert oracle.validate_ragged(3, np.array(rs, np.int64)) == code
    with pytest.raises(oracle.OracleError) as e:
        oracle.group_gather(np.arange(3, dtype=np.int64), np.array(rs, np.int64), [0], [0, 1])
    assert e.value.code == oracle.INVALID_ARGUMENT


def test_bitmap_chain_golden(oracle):
    c = KAT["bitmap_ref_difference_chain"]
    flags = np.array(c["flags0"], np.int32)
    for call in c["calls"]:
        v, rs, _ = oracle.bitmap_ref_difference(np.array(call["values"], np.int32), call["row_splits"], flags)
        assert v.tolist() == call["c_values"] and rs.tolist() == call["c_row_splits"]
    assert flags.tolist() == c["flags_final"]


def test_bitmap_bit31_and_void(oracle):
    flags = np.zeros(2, np.int32)
    v, rs, _ = oracle.bitmap_ref_difference(np.array([31, 63, 31, 0], np.int32), [0, 4], flags)
    assert v.tolist() == [31, 63, 0]
    assert flags.view(np.uint32).tolist() == [0x80000001, 0x80000000]   # 1<<31 on int32 (bitmap_ops.cc:229)
    v, rs, _ = oracle.bitmap_ref_difference(np.zeros(0, np.int32), [0], flags)
    assert v.size == 0 and rs.tolist() == [0]


def test_bitmap_random_vs_python(oracle):
    rng = np.random.default_rng(5)
    vals = rng.integers(0, 4000, 5000).astype(np.int64)
    rs = np.array([0, 100, 100, 3000, 5000])
    flags = np.zeros(125, np.int32)
    got, grs, _ = oracle.bitmap_ref_difference(vals, rs, flags)
    seen, want, wrs = set(), [], [0]
    for g in range(4):
        for x in vals[rs[g]:rs[g + 1]]:
            if x not in seen:
                seen.add(int(x)); want.append(int(x))
        wrs.append(len(want))
    assert got.tolist() == want and grs.tolist() == wrs
    bits = np.unpackbits(flags.view(np.uint8), bitorder="little")
    assert set(np.nonzero(bits)[0].tolist()) == seen


def test_topk_golden(oracle):
    for c in KAT["topk_v2"]:
        v, i = oracle.top_k(np.array(c["input"], np.float32), c["k"])
        np.testing.assert_array_equal(i, np.array(c["indices"]), err_msg=c["source"])
        np.testing.assert_allclose(v, np.array(c["values"], np.float32), rtol=0, atol=0, err_msg=c["source"])
    for c in KAT["topk_v2_errors"]:
        with pytest.raises(oracle.OracleError) as e:
            oracle.top_k(np.array(c["input"], np.float32), c["k"])
        assert e.value.code == oracle.INVALID_ARGUMENT


@pytest.mark.parametrize("n,k", [(500, 1), (500, 5), (500, 50), (500, 500), (5000, 4999), (6140, 5), (26000, 400)])
def test_topk_vs_stable_argsort(oracle, n, k):
    # topk_op_test.py testStableSort / testLargeSort / testLargeTopK / testMediumTopK / testTop3
    rng = np.random.default_rng(n + k)
    x = rng.integers(0, 4, (3, n)).astype(np.float32)           # many ties
    y = rng.standard_normal((3, n)).astype(np.float32)
    for inp in (x, y):
        want = np.argsort(-inp, axis=1, kind="stable")[:, :k]
        v, i = oracle.top_k(inp, k)
        np.testing.assert_array_equal(i, want)
        np.testing.assert_array_equal(v, np.take_along_axis(inp, want, 1))


def test_topk_signed_zero_is_a_tie(oracle):
    v, i = oracle.top_k(np.array([-0.0, 0.0, -1.0, 0.0], np.float32), 3)
    assert i.tolist() == [0, 1, 3]


def test_huge_const_golden(oracle, tmp_path):
    for n, c in enumerate(KAT["huge_const"]):
        a = np.array(c["array"], c["dtype"])
        p = str(tmp_path / f"huge{n}.npy")
        np.save(p, a)
        st, got = oracle.huge_const_check(p, a.dtype, a.shape, read=True)
        assert st == oracle.OK
        np.testing.assert_array_equal(got, a)
        other = np.float32 if a.dtype != np.float32 else np.int32
        assert oracle.huge_const_check(p, other, a.shape)[0] == oracle.INTERNAL       # dtype mismatch
        assert oracle.huge_const_check(p, a.dtype, (9,) + a.shape[1:])[0] == oracle.INTERNAL  # shape mismatch
    assert oracle.huge_const_check(str(tmp_path / "missing.npy"), np.float32, (1,))[0] == oracle.NOT_FOUND
    f = np.asfortranarray(np.arange(6, dtype=np.float32).reshape(2, 3))
    np.save(str(tmp_path / "f.npy"), f)
    assert oracle.huge_const_check(str(tmp_path / "f.npy"), np.float32, (2, 3))[0] == oracle.UNIMPLEMENTED


def test_mlp_blocked_equals_definition(oracle, small_world):
    m = oracle.Mlp(*small_world["mlp"])
    emb, u = small_world["emb"], small_world["queries"][0]
    ids = np.random.default_rng(0).integers(0, emb.shape[0], 257).astype(np.int32)
    a = m.score_def(u, emb[ids])
    b = m.score(u, emb, ids)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))          # bit-identical
    W1, b1, W2, b2, w3 = [x.astype(np.float64) for x in small_world["mlp"]]
    x = np.concatenate([np.tile(u, (257, 1)), emb[ids]], 1).astype(np.float64)
    ref = np.maximum(np.maximum(x @ W1.T + b1, 0) @ W2.T + b2, 0) @ w3
    assert np.abs(ref - b).max() < 2e-6


def test_attention_vs_float64(oracle):
    from nann_b200 import scorer_weights as sw
    blob = sw.attention_blob(seed=3)
    a = oracle.Attn(blob)
    rng = np.random.default_rng(1)
    user = (0.01 * rng.random((50, 64))).astype(np.float32)      # gen_runmeta.py:28-29 scale
    items = (rng.standard_normal((40, 64)) / 8).astype(np.float32)
    got = a.score(user, items)
    # independent float64 restatement of model.py:189-233
    p = blob.astype(np.float64); o = [0]
    def take(*shape):
        n = int(np.prod(shape)); r = p[o[0]:o[0] + n].reshape(shape); o[0] += n; return r
    prelu = lambda x, al: np.maximum(x, 0) + al * np.minimum(x, 0)
    Wq1, bq1, aq, Wq2, bq2 = take(64, 128), take(128), take(128), take(128, 256), take(256)
    Wk1, bk1, ak, Wk2, bk2 = take(64, 128), take(128), take(128), take(128, 256), take(256)
    q = prelu(items @ Wq1 + bq1, aq) @ Wq2 + bq2
    k = prelu(user @ Wk1 + bk1, ak) @ Wk2 + bk2
    att = q @ k.T / 16.0
    att = np.exp(att - att.max(1, keepdims=True)); att /= att.sum(1, keepdims=True)
    h = np.concatenate([att @ user.astype(np.float64), items], 1)
    for i, oo in ((128, 128), (128, 64), (64, 32)):
        W, b, s, t, al = take(i, oo), take(oo), take(oo), take(oo), take(oo)
        h = prelu((h @ W + b) * s + t, al)
    ref = h @ take(32)
    assert np.abs(ref - got).max() < 1e-5


def test_search_stats_and_shapes(oracle, small_world):
    from tests import util
    ix = util.oracle_index(oracle, small_world)
    m = oracle.Mlp(*small_world["mlp"])
    u = small_world["queries"][3]
    T = small_world["T"]
    res = ix.search(lambda r, ids: m.score(u, small_world["emb"], ids), T, trace=True)
    assert res["status"] == oracle.OK
    assert res["n_scored"][0] == len(small_world["ep"])
    assert len(set(res["nodes"].tolist())) == T[5]                       # visited set => no duplicates
    assert np.all(np.diff(res["scores"]) <= 0)
    np.testing.assert_array_equal(res["ids"], small_world["item_ids"][res["nodes"]])
    for r in range(5):
        ids, sc = res["trace"][r]
        assert len(ids) == res["n_scored"][r] and len(set(ids.tolist())) == len(ids)
    # batch helper gives the same answer as the callback form
    b = ix.search_batch_mlp(m, small_world["queries"][:8], T, nthreads=2)
    np.testing.assert_array_equal(b["ids"][3], res["ids"])
    np.testing.assert_array_equal(b["scores"][3].view(np.uint32), res["scores"].view(np.uint32))


def _exec_pb_numpy(w, score, T):
    """build_opt_graph.py:109-149 written with numpy primitives only (independent of the C oracle): ragged gather =
    concatenation of CSR rows in frontier order, set_difference = first occurrence not yet flagged, top_k = stable
    descending argsort (TopKV2's tie rule)."""
    nbr = [(w["values"][l], w["row_splits"][l]) for l in range(2)]

    def expand(level, ids):
        v, rs = nbr[level]
        return np.concatenate([v[rs[i]:rs[i + 1]] for i in ids]) if len(ids) else np.zeros(0, np.int64)

    def diff(ids, flags):
        out = []
        for v in ids:                                   # bitmap_ops.cc:221-234: keep iff bit unset, then set it
            if not flags[v]:
                flags[v] = True
                out.append(v)
        return np.asarray(out, np.int64)

    def topk(ids, sc, k):
        assert len(ids) >= k
        o = np.argsort(-sc, kind="stable")[:k]
        return ids[o], sc[o]

    n = w["emb"].shape[0]
    ep = w["ep"].astype(np.int64)
    R, Rs = topk(ep, score(ep), T[0])                                    # level 2   :109-112
    N1 = expand(1, R)                                                    # level 1   :114-127
    flags = np.zeros(n, bool)
    R = diff(R, flags)
    N1 = diff(N1, flags)
    s1 = score(N1)
    R, Rs = topk(np.concatenate([R, N1]), np.concatenate([Rs, s1]), T[1])
    flags = np.zeros(n, bool)                                            # level 0   :128-141 (visited set reset)
    C = diff(R.copy(), flags)
    for i in range(3):
        Nx = diff(expand(0, C), flags)
        sx = score(Nx)
        C, Cs = topk(Nx, sx, T[i + 2])
        R, Rs = np.concatenate([R, C]), np.concatenate([Rs, Cs])
    R, Rs = topk(R, Rs, T[5])                                            # :143-144
    return w["item_ids"][R], Rs


def test_search_matches_numpy_restatement(oracle, small_world):
    """Pins orc_search (the whole exec.pb dataflow, not just its ops) against an independent numpy restatement, with
    the mlp scorer and with a scorer full of ties (scores rounded to one decimal: exercises both tie rules --
    position in the concatenated list, old results before new ones)."""
    from tests import util
    w = small_world
    ix = util.oracle_index(oracle, w)
    m = oracle.Mlp(*w["mlp"])
    for T in (w["T"], [30, 45, 20, 70, 25, 60]):
        for q in (1, 9):
            u = w["queries"][q]
            for tie in (False, True):
                def score(ids, u=u, tie=tie):
                    s = m.score(u, w["emb"], np.asarray(ids, np.int32)) if len(ids) else np.zeros(0, np.float32)
                    return np.round(s, 1).astype(np.float32) if tie else s
                want_ids, want_sc = _exec_pb_numpy(w, score, T)
                got = ix.search(lambda r, ids: score(ids), T)
                assert got["status"] == oracle.OK
                np.testing.assert_array_equal(got["ids"], want_ids)
                np.testing.assert_array_equal(got["scores"].view(np.uint32), want_sc.view(np.uint32))


def test_search_error_when_fewer_than_k(oracle, small_world):
    from tests import util
    ix = util.oracle_index(oracle, small_world)
    m = oracle.Mlp(*small_world["mlp"])
    u = small_world["queries"][0]
    T = [len(small_world["ep"]) + 1, 10, 10, 10, 10, 10]                 # topk_op.cc:66-69
    res = ix.search(lambda r, ids: m.score(u, small_world["emb"], ids), T)
    assert res["status"] == oracle.INVALID_ARGUMENT


def test_batch_topk_on_rt_golden(oracle):
    c = KAT["batch_topk_on_rt"]
    for call in c["calls"]:
        v, i, rs = oracle.batch_top_k_on_rt(c["values"], c["row_splits"], call["k"], call["ascending"])
        assert v.tolist() == call["values_out"] and i.tolist() == call["idx_out"] and rs.tolist() == call["row_splits_out"]
    v, i, rs = oracle.batch_top_k_on_rt(c["void"]["values"], c["void"]["row_splits"], c["void"]["k"])
    assert v.size == 0 and i.size == 0 and rs.tolist() == [0]
    with pytest.raises(oracle.OracleError):
        oracle.batch_top_k_on_rt(c["values"], c["row_splits"], [1, 2])          # k vector length != groups


def test_fingerprint64_golden(oracle):
    """farmhash::Fingerprint64 restatement against the values the reference's own tests state"""
    for rec in KAT["fingerprint64"]:
        for s_, want in zip(rec["inputs"], rec["values"]):
            assert oracle.fingerprint64(s_.encode()) == want, (rec["source"], s_)
        if "mod10" in rec:
            assert [oracle.fingerprint64(s_.encode()) % 10 for s_ in rec["inputs"]] == rec["mod10"]
    assert oracle.fingerprint64(b"") == 0x9ae16a3b2f90404f


def _py_bloom(values, row_splits, flags, bucket, bucket_size, fp):
    """independent restatement of bitmap_ops.cc:334-359 in Python integers"""
    import math
    primes = []
    for mp in (29, 47, 67, 83):
        t = mp * bucket_size * 32
        while not all(t % i for i in range(2, int(math.sqrt(t) + 1e-6) + 1)):
            t -= 1
        primes.append(t)
    out, rs = [], [0]
    for g in range(len(row_splits) - 1):
        for j in range(row_splits[g], row_splits[g + 1]):
            raw = fp(str(int(values[j])).encode())
            if bucket > 0:
                raw %= bucket
            miss = 0
            for l, mult in enumerate((1, 3, 5, 7)):
                b = (((raw * mult) & ((1 << 64) - 1)) % primes[l]) % (bucket_size * 32)
                w, bit = b >> 5, b & 31
                if not (int(flags[w]) & 0xFFFFFFFF) >> bit & 1:
                    miss += 1
                    flags[w] = np.int32(np.uint32((int(flags[w]) & 0xFFFFFFFF) | (1 << bit)).view(np.int32)) if False else np.array([(int(flags[w]) & 0xFFFFFFFF) | (1 << bit)], np.uint32).view(np.int32)[0]
            if miss:
                out.append(int(values[j]))
        rs.append(len(out))
    return out, rs, primes


def test_bloom_filter_difference_chain_golden(oracle):
    rec = KAT["bloom_filter_difference_chain"]
    flags = np.array(rec["flags0"], np.int32)
    pflags = flags.copy()
    for call in rec["calls"]:
        c, crs, _ = oracle.bloom_filter_difference(np.array(call["values"], np.int32), call["row_splits"], flags,
                                                   rec["bucket"], rec["bucket_size"])
        assert c.tolist() == call["c_values"] and crs.tolist() == call["c_row_splits"]
        pc, prs, primes = _py_bloom(call["values"], call["row_splits"], pflags, rec["bucket"], rec["bucket_size"], oracle.fingerprint64)
        assert pc == call["c_values"] and prs == call["c_row_splits"] and primes == rec["primes"]
    assert flags.tolist() == rec["flags_final"] and pflags.tolist() == rec["flags_final"]


def test_bloom_filter_difference_random_vs_python(oracle):
    """collisions, duplicates, int64 ids beyond 16 digits (the 17..32-byte hash branch), bucket > 0, void and bad input"""
    rng = np.random.default_rng(3)
    for dt, hi in ((np.int32, 2000), (np.int64, 1 << 62)):
        flags = np.zeros(4, np.int32)                              # 128 bits: heavy collisions
        pflags = flags.copy()
        for _ in range(4):
            vals = rng.integers(-5 if dt == np.int64 else 0, hi, 60).astype(dt)
            vals[10:20] = vals[:10]
            rs = [0, 25, 25, 60]
            c, crs, _ = oracle.bloom_filter_difference(vals, rs, flags, 1000003, 4)
            pc, prs, _ = _py_bloom(vals.tolist(), rs, pflags, 1000003, 4, oracle.fingerprint64)
            assert c.tolist() == pc and crs.tolist() == prs
            np.testing.assert_array_equal(flags, pflags)
    c, crs, _ = oracle.bloom_filter_difference(np.zeros(0, np.int32), [0], np.zeros(2, np.int32), 0, 2)
    assert c.size == 0 and crs.tolist() == [0]
    with pytest.raises(oracle.OracleError):
        oracle.bloom_filter_difference(np.array([1, 2], np.int32), [0, 3], np.zeros(2, np.int32), 0, 2)
