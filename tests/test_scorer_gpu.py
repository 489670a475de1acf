"""Scorer parity.  EXACT mlp path: bit-identical to the oracle's fp32 definition.  Attention
scorer and TENSOR mlp path: |diff| <= 1e-5 (north_star tolerance)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def nb():
    import nann_b200
    return nann_b200


@pytest.mark.parametrize("n", [2, 63, 64, 65, 1000, 14800])
def test_mlp_exact_bit_identical(nb, oracle, small_world, n):
    emb = small_world["emb"]
    m = oracle.Mlp(*small_world["mlp"])
    s = nb.Scorer.mlp(*small_world["mlp"])
    rng = np.random.default_rng(n)
    ids = rng.integers(0, emb.shape[0], n).astype(np.int32)
    u = small_world["queries"][n % 64]
    want = m.score(u, emb, ids)
    got = nb.score_ids(s, u, emb, ids)                       # fused gather + score
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
    got2 = nb.blaze_xla_op(s, u, emb[ids])                   # BlazeXlaOp form: dense item_emb input
    np.testing.assert_array_equal(got2.view(np.uint32), want.view(np.uint32))


def test_mlp_exact_adversarial_values(nb, oracle, small_world):
    """denormals, large magnitudes, exact zeros: still bit-identical (no FTZ, IEEE fma)."""
    W1, b1, W2, b2, w3 = [a.copy() for a in small_world["mlp"]]
    rng = np.random.default_rng(9)
    x = rng.standard_normal((300, 128)).astype(np.float32)
    x[:50] *= 1e-38
    x[50:100] *= 1e4
    x[100:110] = 0
    W1[::7] *= 1e-20
    m = oracle.Mlp(W1, b1, W2, b2, w3)
    s = nb.Scorer.mlp(W1, b1, W2, b2, w3)
    u = (rng.standard_normal(128) * 3).astype(np.float32)
    np.testing.assert_array_equal(nb.blaze_xla_op(s, u, x).view(np.uint32), m.score(u, x).view(np.uint32))


def test_attention_scorer(nb, oracle):
    from nann_b200 import scorer_weights as sw
    blob = sw.attention_blob(seed=3)
    a = oracle.Attn(blob)
    s = nb.Scorer.attention(blob)
    assert s.user_floats == 3200 and s.item_dim == 64
    rng = np.random.default_rng(1)
    user = (0.01 * rng.random((50, 64))).astype(np.float32)          # gen_runmeta.py:28-29
    table = (rng.standard_normal((4000, 64)) / 8).astype(np.float16).astype(np.float32)   # f16-representable rows
    for n in (2, 31, 32, 33, 1000):
        ids = rng.integers(0, 4000, n).astype(np.int32)
        want = a.score(user, table, ids)
        got = nb.score_ids(s, user, table, ids)
        assert np.abs(got - want).max() <= TOL
        got2 = nb.blaze_xla_op(s, user, table[ids])
        assert np.abs(got2 - want).max() <= TOL
    user2 = rng.standard_normal((50, 64)).astype(np.float32)          # O(1) inputs: sharper softmax
    want = a.score(user2, table, ids)
    assert np.abs(nb.score_ids(s, user2, table, ids) - want).max() <= TOL * max(1.0, np.abs(want).max())


def test_scorer_id_range_check(nb, small_world):
    s = nb.Scorer.mlp(*small_world["mlp"])
    with pytest.raises(nb.NannError) as e:
        nb.score_ids(s, small_world["queries"][0], small_world["emb"][:100], np.array([1, 100], np.int32))
    assert e.value.code == nb._lib.INVALID_ARGUMENT


@pytest.mark.parametrize("n", [2, 127, 128, 129, 1000, 14800, 60000])
def test_mlp_tensor_core_within_tolerance(nb, oracle, small_world, n):
    """NANN_SCORER_TENSOR (tcgen05, fp16 hi/lo split, fp32 accumulate): |score - oracle| <= 1e-5."""
    emb = small_world["emb"]
    m = oracle.Mlp(*small_world["mlp"])
    s = nb.Scorer.mlp(*small_world["mlp"])
    s.set_precision(nb.SCORER_TENSOR)
    rng = np.random.default_rng(n)
    ids = rng.integers(0, emb.shape[0], n).astype(np.int32)
    u = small_world["queries"][n % 64]
    want = m.score(u, emb, ids)
    got = nb.score_ids(s, u, emb, ids)
    assert not np.isnan(got).any()
    assert np.abs(got - want).max() <= TOL
    got2 = nb.blaze_xla_op(s, u, emb[ids])
    assert np.abs(got2 - want).max() <= TOL
    s.set_precision(nb.SCORER_EXACT)                     # switching back restores bit-exactness
    np.testing.assert_array_equal(nb.score_ids(s, u, emb, ids).view(np.uint32), want.view(np.uint32))


def test_mlp_tensor_core_larger_magnitudes(nb, oracle, small_world):
    """inputs 30x larger than the synthetic corpus: the error bound is relative to the score scale."""
    W1, b1, W2, b2, w3 = small_world["mlp"]
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((2000, 128)) * 3).astype(np.float32)
    u = (rng.standard_normal(128) * 3).astype(np.float32)
    m = oracle.Mlp(W1, b1, W2, b2, w3)
    s = nb.Scorer.mlp(W1, b1, W2, b2, w3)
    s.set_precision(nb.SCORER_TENSOR)
    want, got = m.score(u, x), nb.blaze_xla_op(s, u, x)
    assert np.abs(got - want).max() <= 1e-5 * max(1.0, float(np.abs(want).max()))


def test_blaze_xla_op_admission_control(nb):
    """BlazeXlaOp::Schedule (blaze_xla_kernel.cc:221-258): running cap, 'waiting pool is full', 'blaze wait too long'."""
    import threading
    from nann_b200 import index as nix, scorer_weights as sw
    sc = nb.Scorer.mlp(*sw.mlp_weights())
    st = sc.admission_state()
    assert st["running_max"] == int(os.environ.get("BLAZE_THREADS_NUM", 2)) and st["max_waiting"] == int(os.environ.get("DENSE_MAX_WAITING_COUNT", 10))
    emb = nix.synthetic_corpus(1_000_000, 128, seed=1)                   # EXACT scorer: ~17 ms per run
    import torch
    emb_d = torch.from_numpy(emb).cuda()
    user = emb[0]

    def fire(n_threads, results):
        def work(i):
            try:
                nb.blaze_xla_op(sc, user, emb_d)
                results[i] = "ok"
            except nb.NannError as e:
                results[i] = (e.code, e.message)
        th = [threading.Thread(target=work, args=(i,)) for i in range(n_threads)]
        [t.start() for t in th]
        [t.join() for t in th]

    # (a) one runner, nobody may wait: concurrent calls beyond the first fail with Internal "waiting pool is full"
    sc.set_admission(running_max=1, max_waiting=0, wait_ms=0)
    res = [None] * 6
    fire(6, res)
    assert res.count("ok") >= 1 and any(r != "ok" for r in res)
    assert all(r == "ok" or (r[0] == 13 and "waiting pool is full" in r[1]) for r in res)
    # (b) everybody may wait: all succeed, never more than one running
    sc.set_admission(running_max=1, max_waiting=10, wait_ms=0)
    res = [None] * 6
    fire(6, res)
    assert res == ["ok"] * 6
    # (c) a 1-ms deadline against runs that take far longer: late calls fail with "blaze wait too long"
    sc.set_admission(running_max=1, max_waiting=10, wait_ms=1)
    res = [None] * 6
    fire(6, res)
    assert res.count("ok") >= 1 and any(r != "ok" for r in res)
    assert all(r == "ok" or (r[0] in (4, 13) and "blaze wait too long" in r[1]) for r in res)
    s2 = sc.admission_state()
    assert s2["running"] == 0 and s2["waiting"] == 0
