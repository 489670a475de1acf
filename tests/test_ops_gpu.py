"""Op-level parity on the GPU: every C-ABI op against the reference's vectors (tests/golden) and
against the CPU oracle on seeded random inputs.  Integer/index results are compared bit-exactly."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat_ops.json")))


@pytest.fixture(scope="module")
def nb():
    import nann_b200
    info = nann_b200.device_info(0)
    assert info["cc"][0] == 10, info
    return nann_b200


def test_group_gather_golden(nb):
    for c in KAT["group_gather"]:
        dt = np.dtype(c["dtype"])
        v, rs = nb.group_gather(np.array(c["params_values"], dt), np.array(c["params_row_splits"], np.int64),
                                np.array(c["indices_values"], np.int64), np.array(c["indices_row_splits"], np.int64),
                                unique=c["unique"])
        assert rs.tolist() == c["ret_row_splits"], c["source"]
        if "ret_values" in c:
            assert v.tolist() == c["ret_values"], c["source"]
        else:
            for g, want in enumerate(c["ret_sets"]):
                got = v[rs[g]:rs[g + 1]].tolist()
                assert sorted(got) == want and len(set(got)) == len(got)


@pytest.mark.parametrize("rs,code", [([1, 3], 2), ([0, 2], 3)])
def test_ragged_validation_codes(nb, rs, code):
    with pytest.raises(nb.NannError) as e:
        nb.group_gather(np.arange(3, dtype=np.int64), np.array(rs, np.int64), np.array([0]), np.array([0, 1]))
    assert e.value.code == nb._lib.INVALID_ARGUMENT and f"code: {code}" in e.value.message
    with pytest.raises(nb.NannError) as e:
        nb.bitmap_ref_difference(np.arange(3, dtype=np.int32), np.array(rs, np.int64), np.zeros(4, np.int32))
    assert e.value.code == nb._lib.INVALID_ARGUMENT and f"code: {code}" in e.value.message


@pytest.mark.parametrize("dt", [np.int32, np.int64])
def test_group_gather_random_vs_oracle(nb, oracle, dt):
    rng = np.random.default_rng(7)
    lens = rng.integers(0, 70, 3000)
    prs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    pv = rng.integers(0, 1 << 20, prs[-1]).astype(dt)
    glen = rng.integers(0, 400, 9)
    irs = np.concatenate([[0], np.cumsum(glen)]).astype(np.int64)
    iv = rng.integers(0, 3000, irs[-1]).astype(np.int64)
    want = oracle.group_gather(pv, prs, iv, irs)
    got = nb.group_gather(pv, prs, iv, irs)
    np.testing.assert_array_equal(got[0], want[0])
    np.testing.assert_array_equal(got[1], want[1])
    cut = int(prs[40])                                  # a valid ragged prefix: rows 0..39
    sub_pv, sub_prs = pv[:cut] % 97, prs[:41]
    sub_iv, sub_irs = iv[:50] % 40, np.array([0, 20, 50], np.int64)
    wu = oracle.group_gather(sub_pv, sub_prs, sub_iv, sub_irs, unique=True)
    gu = nb.group_gather(sub_pv, sub_prs, sub_iv, sub_irs, unique=True)
    np.testing.assert_array_equal(gu[0], wu[0])       # first-occurrence order on both sides
    np.testing.assert_array_equal(gu[1], wu[1])


def test_group_gather_device_inputs(nb, oracle):
    import torch
    rng = np.random.default_rng(3)
    lens = rng.integers(0, 64, 500)
    prs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    pv = rng.integers(0, 500, prs[-1]).astype(np.int32)
    iv = rng.integers(0, 500, 200).astype(np.int64)
    irs = np.array([0, 200], np.int64)
    got = nb.group_gather(torch.from_numpy(pv).cuda(), torch.from_numpy(prs).cuda(), iv, irs)
    want = oracle.group_gather(pv, prs, iv, irs)
    np.testing.assert_array_equal(got[0], want[0])


def test_bitmap_chain_golden(nb):
    c = KAT["bitmap_ref_difference_chain"]
    flags = np.array(c["flags0"], np.int32)
    for call in c["calls"]:
        v, rs, fl = nb.bitmap_ref_difference(np.array(call["values"], np.int32), np.array(call["row_splits"], np.int64), flags)
        assert v.tolist() == call["c_values"] and rs.tolist() == call["c_row_splits"]
        assert fl is flags
    assert flags.tolist() == c["flags_final"]


@pytest.mark.parametrize("dt", [np.int32, np.int64])
def test_bitmap_random_vs_oracle(nb, oracle, dt):
    rng = np.random.default_rng(11)
    n_items = 100_000
    vals = np.concatenate([rng.integers(0, n_items, 20000), rng.integers(0, 50, 3000), [n_items - 1, 31, 31, 0]]).astype(dt)
    rs = np.array([0, 17, 17, 9000, 20000, vals.size], np.int64)
    f_gpu = np.zeros((n_items + 31) // 32, np.int32)
    f_cpu = f_gpu.copy()
    for _ in range(2):   # second pass runs against the already-populated bitmap
        got = nb.bitmap_ref_difference(vals, rs, f_gpu)
        want = oracle.bitmap_ref_difference(vals, rs, f_cpu)
        np.testing.assert_array_equal(got[0], want[0])
        np.testing.assert_array_equal(got[1], want[1])
        np.testing.assert_array_equal(f_gpu, f_cpu)
    v, r, _ = nb.bitmap_ref_difference(np.zeros(0, dt), np.array([0], np.int64), f_gpu)   # void input
    assert v.size == 0 and r.tolist() == [0]


def test_bitmap_out_of_range_is_an_error(nb):
    with pytest.raises(nb.NannError) as e:
        nb.bitmap_ref_difference(np.array([5, 64], np.int32), np.array([0, 2], np.int64), np.zeros(2, np.int32))
    assert e.value.code == nb._lib.INVALID_ARGUMENT


def test_bitmap_device_flags_in_place(nb, oracle):
    import torch
    flags = torch.zeros(8, dtype=torch.int32, device="cuda")
    v, rs, _ = nb.bitmap_ref_difference(np.array([3, 3, 200, 3, 255], np.int32), np.array([0, 5], np.int64), flags)
    assert v.tolist() == [3, 200, 255]
    f_cpu = np.zeros(8, np.int32)
    oracle.bitmap_ref_difference(np.array([3, 3, 200, 3, 255], np.int32), [0, 5], f_cpu)
    np.testing.assert_array_equal(flags.cpu().numpy(), f_cpu)


def test_topk_golden(nb):
    for c in KAT["topk_v2"]:
        v, i = nb.top_k(np.array(c["input"], np.float32), c["k"])
        np.testing.assert_array_equal(i, np.array(c["indices"]), err_msg=c["source"])
        np.testing.assert_array_equal(v, np.array(c["values"], np.float32), err_msg=c["source"])
    for c in KAT["topk_v2_errors"]:
        with pytest.raises(nb.NannError) as e:
            nb.top_k(np.array(c["input"], np.float32), c["k"])
        assert e.value.code == nb._lib.INVALID_ARGUMENT and c["message"] in e.value.message
    v, i = nb.top_k(np.zeros((0, 10), np.float32), 3)                  # topk_op_test.py testTop3ZeroRows
    assert v.shape == (0, 3) and i.shape == (0, 3)
    v, i = nb.top_k(np.arange(5, dtype=np.float32), 0)
    assert v.shape == (0,)


@pytest.mark.parametrize("n,k", [(2, 1), (33, 33), (500, 1), (500, 5), (500, 50), (500, 500), (5000, 4096),
                                 (6140, 5), (26000, 400), (100000, 1000)])
def test_topk_vs_oracle(nb, oracle, n, k):
    rng = np.random.default_rng(n * 7 + k)
    cases = [rng.integers(0, 4, (3, n)).astype(np.float32),                         # heavy ties (testStableSort)
             rng.standard_normal((3, n)).astype(np.float32),
             np.where(rng.random((2, n)) < 0.5, -0.0, 0.0).astype(np.float32),       # signed zeros tie
             (rng.standard_normal((2, n)) * 1e-42).astype(np.float32)]               # denormals
    for inp in cases:
        wv, wi = oracle.top_k(inp, k)
        gv, gi = nb.top_k(inp, k)
        np.testing.assert_array_equal(gi, wi)
        np.testing.assert_array_equal(gv.view(np.uint32), wv.view(np.uint32))


def test_topk_full_size_properties(nb):
    """size-independent checks at BASELINE sizes: 1024 rows x 25600 cols, k=400."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(1024, 25600, device="cuda", generator=g)
    v, i = nb.top_k(x, 400)
    xs = x.cpu().numpy()
    assert np.all(np.diff(v, axis=1) <= 0)                                           # sorted
    np.testing.assert_array_equal(np.take_along_axis(xs, i.astype(np.int64), 1), v)  # values match indices
    kth = v[:, -1:]
    assert np.all((xs > kth).sum(1) <= 399)                                          # nothing better left out
    assert all(len(set(r.tolist())) == 400 for r in i[:64])


def test_gather_rows(nb, oracle):
    rng = np.random.default_rng(2)
    table = rng.standard_normal((5000, 128)).astype(np.float32)
    ids = rng.integers(0, 5000, 3333).astype(np.int32)
    np.testing.assert_array_equal(nb.gather(table, ids), table[ids])
    t16 = table[:, :64].astype(np.float16)
    np.testing.assert_array_equal(nb.gather(np.ascontiguousarray(t16), ids), t16[ids])
    i64 = rng.integers(0, 1 << 40, 5000).astype(np.int64)
    np.testing.assert_array_equal(nb.gather(i64, ids), i64[ids])                     # item_ids gather (:144)
    odd = rng.integers(0, 255, (100, 7)).astype(np.uint8)
    np.testing.assert_array_equal(nb.gather(odd, ids % 100), odd[ids % 100])
    with pytest.raises(nb.NannError) as e:
        nb.gather(table, np.array([0, 5000], np.int32))
    assert e.value.code == nb._lib.INVALID_ARGUMENT


def test_huge_const_to_device(nb, tmp_path):
    import torch
    a = np.random.default_rng(0).standard_normal((1000, 128)).astype(np.float32)
    p = str(tmp_path / "t.npy")
    np.save(p, a)
    h = nb.huge_const(p, np.float32, a.shape, device=0)
    assert h.device_ptr
    ids = np.array([5, 999, 0], np.int32)
    out = np.empty((3, 128), np.float32)
    import ctypes as C
    nb._lib.check(nb._lib.lib().nann_gather_rows(C.c_void_p(h.device_ptr), 1000, 512, C.c_void_p(ids.ctypes.data), 3,
                                                 C.c_void_p(out.ctypes.data), None))
    np.testing.assert_array_equal(out, a[ids])
    for c in KAT["huge_const"]:
        arr = np.array(c["array"], c["dtype"])
        q = str(tmp_path / "k.npy")
        np.save(q, arr)
        np.testing.assert_array_equal(nb.huge_const(q, arr.dtype, arr.shape, device=0).numpy(), arr)


def test_merge_topk(nb, oracle):
    rng = np.random.default_rng(4)
    G, B, kin, kout = 8, 37, 200, 200
    sc = -np.sort(-rng.integers(0, 50, (G, B, kin)).astype(np.float32), axis=2)      # per-shard sorted, many ties
    ids = rng.integers(0, 1 << 40, (G, B, kin)).astype(np.int64)
    gs, gi = nb.merge_topk(sc, ids, kout)
    cat_s = sc.transpose(1, 0, 2).reshape(B, G * kin)
    cat_i = ids.transpose(1, 0, 2).reshape(B, G * kin)
    wv, wi = oracle.top_k(cat_s, kout)                                              # concat order = (shard, rank)
    np.testing.assert_array_equal(gs, wv)
    np.testing.assert_array_equal(gi, np.take_along_axis(cat_i, wi.astype(np.int64), 1))
    with pytest.raises(nb.NannError):
        nb.merge_topk(sc[:, :, :10], ids[:, :, :10], 200)


def test_batch_topk_on_rt(nb, oracle):
    c = KAT["batch_topk_on_rt"]
    for call in c["calls"]:
        v, i, rs = nb.batch_top_k_on_rt(np.array(c["values"], np.float32), np.array(c["row_splits"], np.int64),
                                        call["k"], call["ascending"])
        assert v.tolist() == call["values_out"] and i.tolist() == call["idx_out"] and rs.tolist() == call["row_splits_out"]
    v, i, rs = nb.batch_top_k_on_rt(np.zeros(0, np.float32), np.array([0], np.int64), 3)
    assert v.size == 0 and i.size == 0 and rs.tolist() == [0]
    with pytest.raises(nb.NannError) as e:
        nb.batch_top_k_on_rt(np.array(c["values"], np.float32), np.array(c["row_splits"], np.int64), [1, 2])
    assert e.value.code == nb._lib.INVALID_ARGUMENT
    rng = np.random.default_rng(12)                                              # a batch of queries' candidate lists
    lens = rng.integers(0, 3000, 40)
    rs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    vals = rng.integers(0, 200, rs[-1]).astype(np.float32)                        # ties
    ks = rng.integers(0, 500, 40).astype(np.int64)
    for asc in (False, True):
        want = oracle.batch_top_k_on_rt(vals, rs, ks, asc)
        got = nb.batch_top_k_on_rt(vals, rs, ks, asc)
        for w, g in zip(want, got):
            np.testing.assert_array_equal(g, w)


def test_ragged_batch_helpers(nb):
    """BatchGatherOnRT / BatchConcatOnRT / SplitsGather / BitmapInit / BitmapDifference against the values the
    reference's own scripts state ("should be ...") and against numpy restatements of the kernels."""
    from nann_b200 import ops
    i64 = lambda a: np.array(a, np.int64)
    c = KAT["batch_gather_on_rt"]
    v, rs = ops.batch_gather_on_rt(i64(c["params_values"]), i64(c["params_row_splits"]), i64(c["indices_values"]), i64(c["indices_row_splits"]))
    assert v.tolist() == c["ret_values"] and rs.tolist() == c["ret_row_splits"]
    for pv, prs, iv, irs in ((c["params_values"], c["params_row_splits"], [], [0]), ([], [0], c["indices_values"], c["indices_row_splits"]), ([], [0], [], [0])):
        v, rs = ops.batch_gather_on_rt(i64(pv), i64(prs), i64(iv), i64(irs))     # "should be []"
        assert v.size == 0 and rs.tolist() == [0]
    with pytest.raises(nb.NannError):
        ops.batch_gather_on_rt(i64([1, 2, 3]), i64([0, 3]), i64([0, 1, 1]), i64([0, 2, 3]))   # row_splits do NOT match
    c = KAT["batch_concat_on_rt"]
    v, rs = ops.batch_concat_on_rt(i64(c["left_values"]), i64(c["left_row_splits"]), i64(c["right_values"]), i64(c["right_row_splits"]))
    assert v.tolist() == c["ret_values"] and rs.tolist() == c["ret_row_splits"]
    v, rs = ops.batch_concat_on_rt(i64(c["left_values"]), i64(c["left_row_splits"]), i64([]), i64([0]))
    assert v.tolist() == c["left_values"] and rs.tolist() == c["left_row_splits"]             # "should be [[1,2,3],[4,5]]"
    v, rs = ops.batch_concat_on_rt(i64([]), i64([0]), i64(c["right_values"]), i64(c["right_row_splits"]))
    assert v.tolist() == c["right_values"] and rs.tolist() == c["right_row_splits"]           # "should be [[0,1],[1]]"
    c = KAT["splits_gather"]
    v, rs = ops.splits_gather(i64(c["splits"]), i64(c["indices_values"]), i64(c["indices_row_splits"]))
    assert v.tolist() == c["ret_values"] and rs.tolist() == c["ret_row_splits"]
    v, rs = ops.splits_gather(i64([]), i64(c["indices_values"]), i64(c["indices_row_splits"]))
    assert v.size == 0 and rs.tolist() == [0]
    # random: numpy restatements
    rng = np.random.default_rng(8)
    lens = rng.integers(0, 50, 200); prs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    pv = rng.integers(0, 1000, prs[-1]).astype(np.int32)
    ilen = np.where(lens > 0, rng.integers(0, 30, 200), 0); irs = np.concatenate([[0], np.cumsum(ilen)]).astype(np.int64)
    iv = np.concatenate([rng.integers(0, max(l, 1), k) for l, k in zip(lens, ilen)]).astype(np.int64)
    v, rs = ops.batch_gather_on_rt(pv, prs, iv, irs)
    want = np.concatenate([pv[prs[g] + iv[irs[g]:irs[g + 1]]] for g in range(200)])
    np.testing.assert_array_equal(v, want); np.testing.assert_array_equal(rs, irs)
    rv = rng.integers(0, 1000, irs[-1]).astype(np.int32)
    v, rs = ops.batch_concat_on_rt(pv, prs, rv, irs)
    want = np.concatenate([np.concatenate([pv[prs[g]:prs[g + 1]], rv[irs[g]:irs[g + 1]]]) for g in range(200)])
    np.testing.assert_array_equal(v, want); np.testing.assert_array_equal(rs, prs + irs)
    bm = ops.bitmap_init(np.array([1, 2, 3, 4, 5, 6, 7, 8, 10, 11, 12, 13, 14], np.int32), 16)
    assert bm.tolist()[:4] == [32254, 0, 0, 0]                                                # same bits as the chained KAT
    flags = np.array([32254, 0, 0, 0], np.int32)
    kept, new = ops.bitmap_difference(np.array([4, 9, 9, 40, 5, 127], np.int32), flags)
    assert kept.tolist() == [9, 40, 127] and flags.tolist() == [32254, 0, 0, 0]              # value semantics: input untouched
    assert new.view(np.uint32).tolist() == [32254 | (1 << 9), 1 << 8, 0, 1 << 31]


def test_bloom_filter_difference(nb, oracle):
    """BloomFilterDifference (bitmap_ops.cc:264-432): the reference script's chain (tests/golden), then random ragged
    batches with heavy collisions / duplicates / negative and > 16-digit int64 ids / bucket > 0, against the oracle:
    values, row_splits and the mutated flags must be identical."""
    import json, os
    rec = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat_ops.json")))["bloom_filter_difference_chain"]
    flags = np.array(rec["flags0"], np.int32)
    for call in rec["calls"]:
        c, crs, _ = nb.bloom_filter_difference(np.array(call["values"], np.int32), np.array(call["row_splits"], np.int64), flags,
                                               bucket=rec["bucket"], bucket_size=rec["bucket_size"])
        assert c.tolist() == call["c_values"] and crs.tolist() == call["c_row_splits"]
    assert flags.tolist() == rec["flags_final"]
    rng = np.random.default_rng(5)
    for dt, lo, hi, bucket, bsz in ((np.int32, 0, 3000, 0, 8), (np.int64, -50, 1 << 62, 1000003, 4), (np.int64, 0, 10 ** 9, 0, 64)):
        g_flags, o_flags = np.zeros(bsz + 3, np.int32), np.zeros(bsz + 3, np.int32)
        for _ in range(4):
            n = int(rng.integers(1, 700))
            vals = rng.integers(lo, hi, n).astype(dt)
            vals[n // 2:] = vals[:n - n // 2]                       # duplicates inside the call
            cuts = np.sort(rng.integers(0, n + 1, 5))
            rs = np.concatenate([[0], cuts, [n]]).astype(np.int64)
            c, crs, _ = nb.bloom_filter_difference(vals, rs, g_flags, bucket=bucket, bucket_size=bsz)
            oc, ors, _ = oracle.bloom_filter_difference(vals, rs, o_flags, bucket, bsz)
            np.testing.assert_array_equal(c, oc)
            np.testing.assert_array_equal(crs, ors)
            np.testing.assert_array_equal(g_flags, o_flags)
    c, crs, _ = nb.bloom_filter_difference(np.zeros(0, np.int32), np.array([0], np.int64), np.zeros(2, np.int32), bucket_size=2)
    assert c.size == 0 and crs.tolist() == [0]                      # void input (:312-322)
    with pytest.raises(nb.NannError) as e:
        nb.bloom_filter_difference(np.array([1, 2], np.int32), np.array([0, 3], np.int64), np.zeros(2, np.int32), bucket_size=2)
    assert e.value.code == 3 and "code: 3" in str(e.value)
    with pytest.raises(nb.NannError):
        nb.bloom_filter_difference(np.array([1], np.int32), np.array([0, 1], np.int64), np.zeros(1, np.int32), bucket_size=2)
