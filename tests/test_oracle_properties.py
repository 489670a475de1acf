"""Property tests of the CPU checker (hypothesis): the size-independent laws of the ops on the path, on ragged, empty and
duplicate-heavy inputs -- the edge cases of the reference's own scripts (empty RaggedTensor: `row_splits = [0]`; groups of
length 0; repeated ids; ties).  The GPU kernels are held to the checker in `-m gpu`; this file holds the checker to the
definitions (GroupGather_kernel.cc:136-170, bitmap_ops.cc:198-257, topk_op.cc:139-207, BatchTopKOnRT_kernel.cc:62-156)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

SET = settings(max_examples=80, deadline=None, derandomize=True, database=None)   # reproducible: same examples on every box


@st.composite
def ragged(draw, max_groups=5, max_len=12, lo=0, hi=60):
    lens = draw(st.lists(st.integers(0, max_len), min_size=0, max_size=max_groups))
    vals = [draw(st.lists(st.integers(lo, hi), min_size=n, max_size=n)) for n in lens]
    rs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    flat = np.array([x for g in vals for x in g], np.int64)
    return vals, flat, rs


@SET
@given(params=ragged(max_groups=8, max_len=6, hi=1000), data=st.data())
def test_group_gather_is_row_concatenation(oracle, params, data):
    pv, pflat, prs = params
    n_rows = len(pv)
    groups = data.draw(st.lists(st.lists(st.integers(0, max(n_rows - 1, 0)), max_size=7), max_size=4)) if n_rows else [[], []]
    iflat = np.array([x for g in groups for x in g], np.int64)
    irs = np.concatenate([[0], np.cumsum([len(g) for g in groups])]).astype(np.int64)
    got, grs = oracle.group_gather(pflat, prs, iflat, irs)
    if n_rows == 0 or len(groups) == 0:         # "void inputs" (GroupGather_kernel.cc:69-77): values [], row_splits [0]
        assert got.size == 0 and grs.tolist() == [0]
        gotu, grsu = oracle.group_gather(pflat, prs, iflat, irs, unique=True)
        assert gotu.size == 0 and grsu.tolist() == [0]
        return
    want = [[x for i in g for x in pv[i]] for g in groups]
    assert got.tolist() == [x for g in want for x in g]
    assert grs.tolist() == np.concatenate([[0], np.cumsum([len(g) for g in want])]).astype(int).tolist()
    # unique=True: first occurrence per group, order kept
    gotu, grsu = oracle.group_gather(pflat, prs, iflat, irs, unique=True)
    wantu = [list(dict.fromkeys(g)) for g in want]
    assert gotu.tolist() == [x for g in wantu for x in g]
    assert grsu.tolist() == np.concatenate([[0], np.cumsum([len(g) for g in wantu])]).astype(int).tolist()


@SET
@given(a=ragged(hi=127), b=ragged(hi=127))
def test_bitmap_ref_difference_laws(oracle, a, b):
    """kept = first occurrence of every id not yet flagged, in order, groups sharing one bitmap; applying the same input
    again keeps nothing (idempotence); the flags are exactly the union of everything seen"""
    flags = np.zeros(4, np.int32)
    seen = set()
    for vals, flat, rs in (a, b, a):
        v, crs, _ = oracle.bitmap_ref_difference(flat.astype(np.int32), rs, flags)
        want, wrs = [], [0]
        for g in vals:
            for x in g:
                if x not in seen:
                    seen.add(x)
                    want.append(x)
            wrs.append(len(want))
        assert v.tolist() == want and crs.tolist() == wrs
    bits = np.unpackbits(flags.view(np.uint8), bitorder="little")
    assert set(np.nonzero(bits)[0].tolist()) == seen


@SET
@given(x=st.lists(st.sampled_from([-2.0, -0.0, 0.0, 0.5, 0.5, 1.0, 3.0, float("inf"), -float("inf")]), min_size=1, max_size=40),
       data=st.data())
def test_topk_is_the_stable_descending_order(oracle, x, data):
    k = data.draw(st.integers(0, len(x)))
    arr = np.array(x, np.float32)
    vals, idx = oracle.top_k(arr, k)
    order = sorted(range(len(x)), key=lambda i: (-arr[i], i))[:k]      # value desc, ties -> lower index; -0.0 == 0.0
    assert idx.tolist() == order
    assert vals.view(np.uint32).tolist() == arr[order].view(np.uint32).tolist()
    with pytest.raises(oracle.OracleError):
        oracle.top_k(arr, len(x) + 1)


@SET
@given(r=ragged(max_groups=6, max_len=10, lo=-5, hi=5), k=st.integers(0, 12), ascending=st.booleans())
def test_batch_topk_on_rt_is_per_group_topk_with_local_indices(oracle, r, k, ascending):
    vals, flat, rs = r
    v, i, ro = oracle.batch_top_k_on_rt(flat.astype(np.float32), rs, k, ascending=ascending)
    wv, wi, wr = [], [], [0]
    for g in vals:
        order = sorted(range(len(g)), key=lambda j: ((g[j] if ascending else -g[j]), j))[:k]
        wv += [float(g[j]) for j in order]
        wi += order
        wr.append(len(wv))
    assert v.tolist() == wv and i.tolist() == wi and ro[:len(wr)].tolist() == wr


@SET
@given(a=ragged(hi=5000), b=ragged(hi=5000))
def test_bloom_filter_difference_never_emits_a_value_twice(oracle, a, b):
    """a Bloom filter has false positives (a new value may be dropped) but no false negatives: nothing is emitted twice,
    whatever the bucket size; output order is input order"""
    for bucket_size in (1, 64):
        flags = np.zeros(4 * bucket_size, np.int32)
        emitted = []
        for vals, flat, rs in (a, b, a):
            v, crs, _ = oracle.bloom_filter_difference(flat.astype(np.int64), rs, flags, bucket=0, bucket_size=bucket_size)
            out = v.tolist()
            assert crs[-1] == len(out) if len(crs) else True
            pos = 0
            for g, (s, e) in zip(vals, zip(crs[:-1], crs[1:])):        # every group's output is a subsequence of its input
                it = iter(g)
                assert all(any(x == y for y in it) for x in out[s:e])
            emitted += out
        assert len(emitted) == len(set(emitted))
