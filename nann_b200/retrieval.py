"""Retrieval = the exec.pb dataflow (NANN_impls/nann/delivery/build_opt_graph.py:69-160).

Two forms, same results:
  * `Searcher.search(users, level_topn)` -- the fused, batched, all-on-device path
    (nann_search_batch): what a server calls.
  * `retrieve_opwise(...)` -- batch=1, one C-ABI call per TF node, mirroring build_model()
    line by line; this is what the TF shim executes when exec.pb runs unchanged.
"""
import ctypes as C
import math

import numpy as np

from . import _lib, ops
from ._lib import check

_NP2CODE = ops._NP2CODE


def _host_or_dev(x, dtype):
    ptr, n, keep = ops._as(x, dtype)
    return ptr, keep


class Index:
    """Appendix-C index files resident in HBM (nann_index_t)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def from_arrays(cls, emb, item_ids, enter_points, nbr_values, nbr_row_splits, device=0):
        """emb f32/f16 [N,d]; item_ids i64 [N]; enter_points i32/i64; nbr_values[l] i32/i64,
        nbr_row_splits[l] i64 [N+1] for l in (0, 1).  numpy arrays or CUDA tensors."""
        if ops._is_torch(emb):
            emb_dt = np.dtype(str(emb.dtype).split(".")[-1])
            n, d = emb.shape
        else:
            emb = np.ascontiguousarray(emb)
            emb_dt = emb.dtype
            n, d = emb.shape
        if emb_dt not in (np.dtype("float32"), np.dtype("float16")):
            raise TypeError("item_embs must be float32 or float16")
        ep_dt = np.dtype(str(enter_points.dtype).split(".")[-1]) if ops._is_torch(enter_points) else np.asarray(enter_points).dtype
        nv_dt = np.dtype(str(nbr_values[0].dtype).split(".")[-1]) if ops._is_torch(nbr_values[0]) else np.asarray(nbr_values[0]).dtype
        e_ptr, k0 = _host_or_dev(emb, emb_dt)
        i_ptr, k1 = _host_or_dev(item_ids, np.int64)
        p_ptr, k2 = _host_or_dev(enter_points, ep_dt)
        v = [_host_or_dev(nbr_values[l], nv_dt) for l in range(2)]
        r = [_host_or_dev(nbr_row_splits[l], np.int64) for l in range(2)]
        n_ep = enter_points.numel() if ops._is_torch(enter_points) else np.asarray(enter_points).size
        vals = (C.c_void_p * 2)(v[0][0], v[1][0])
        rss = (C.c_void_p * 2)(r[0][0], r[1][0])
        n_vals = (C.c_int64 * 2)(*[(x.numel() if ops._is_torch(x) else np.asarray(x).size) for x in nbr_values])
        h = C.c_void_p()
        check(_lib.lib().nann_index_create(n, d, e_ptr, _NP2CODE[emb_dt], i_ptr, p_ptr, _NP2CODE[np.dtype(ep_dt)], n_ep,
                                           vals, n_vals, _NP2CODE[np.dtype(nv_dt)], rss, int(device), C.byref(h)))
        return cls(h)

    @classmethod
    def from_arrays_sharded(cls, n_items, emb_rows, row_lo, item_ids, enter_points, nbr_values, nbr_row_splits, device=0):
        """Index for distributed scoring (nann_index_create_sharded): the graph, enter points and item ids of the WHOLE
        corpus, but only rows [row_lo, row_lo + len(emb_rows)) of the embedding table (numpy arrays)."""
        emb = np.ascontiguousarray(emb_rows)
        if emb.dtype not in (np.dtype("float32"), np.dtype("float16")):
            raise TypeError("item_embs must be float32 or float16")
        ep = np.ascontiguousarray(enter_points)
        vals_np = [np.ascontiguousarray(v) for v in nbr_values]
        rs_np = [np.ascontiguousarray(r, np.int64) for r in nbr_row_splits]
        ids = np.ascontiguousarray(item_ids, np.int64)
        vals = (C.c_void_p * 2)(vals_np[0].ctypes.data, vals_np[1].ctypes.data)
        rss = (C.c_void_p * 2)(rs_np[0].ctypes.data, rs_np[1].ctypes.data)
        n_vals = (C.c_int64 * 2)(vals_np[0].size, vals_np[1].size)
        h = C.c_void_p()
        check(_lib.lib().nann_index_create_sharded(int(n_items), emb.shape[1], C.c_void_p(emb.ctypes.data), _NP2CODE[emb.dtype], int(row_lo),
                                                   emb.shape[0], C.c_void_p(ids.ctypes.data), C.c_void_p(ep.ctypes.data), _NP2CODE[ep.dtype],
                                                   ep.size, vals, n_vals, _NP2CODE[vals_np[0].dtype], rss, int(device), C.byref(h)))
        return cls(h)

    @classmethod
    def load(cls, embs_dir, index_dir, device=0):
        """item_embs.npy / item_ids.npy from embs_dir; enter_points.npy and
        neighbors_level_{0,1}_{values,row_splits}.npy from index_dir (either dtype width)."""
        h = C.c_void_p()
        check(_lib.lib().nann_index_load(str(embs_dir).encode(), str(index_dir).encode(), int(device), C.byref(h)))
        return cls(h)

    n_items = property(lambda self: int(_lib.lib().nann_index_n_items(self._h)))
    dim = property(lambda self: int(_lib.lib().nann_index_dim(self._h)))
    n_enter_points = property(lambda self: int(_lib.lib().nann_index_n_enter_points(self._h)))
    emb_device_ptr = property(lambda self: _lib.lib().nann_index_emb_device(self._h))

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().nann_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class EvalSearcher:
    """The `main.py --job-type test` traversal (NANN_impls/nann/model/model.py:299-362) for batches of users.
    Arguments keep the reference's flag names (nann/config.py:52-57), indexed by level 0..2."""

    def __init__(self, index, scorer, max_batch, max_top_k_per_level=(400, 200, 100), max_topk_eval=200):
        self.index, self.scorer = index, scorer
        self.max_batch = int(max_batch)
        K = (C.c_int32 * 3)(*[int(k) for k in max_top_k_per_level])
        h = C.c_void_p()
        check(_lib.lib().nann_eval_searcher_create(index._h, scorer._h, self.max_batch, K, int(max_topk_eval), C.byref(h)))
        self._h = h

    def search(self, users, num_scoring_per_level=(3, 1, 1), top_k_per_level=(400, 200, 100), topk_eval=200, stream=None):
        """users: [B, user_floats] numpy or CUDA tensor.
        Returns dict(ids i64[B,k] (-1 padded), scores f32[B,k], nodes i32[B,k], n i32[B], status i32[B], n_scored)."""
        uf = self.scorer.user_floats
        if ops._is_torch(users):
            u = users.contiguous().float().reshape(-1, uf)
            B, uptr = u.shape[0], C.c_void_p(u.data_ptr())
        else:
            u = np.ascontiguousarray(users, np.float32).reshape(-1, uf)
            B, uptr = u.shape[0], C.c_void_p(u.ctypes.data)
        k = int(topk_eval)
        ns = (C.c_int32 * 3)(*[int(x) for x in num_scoring_per_level])
        tk = (C.c_int32 * 3)(*[int(x) for x in top_k_per_level])
        ids, sc, nodes = np.empty((B, k), np.int64), np.empty((B, k), np.float32), np.empty((B, k), np.int32)
        n, status = np.empty(B, np.int32), np.empty(B, np.int32)
        tot = C.c_int64(0)
        check(_lib.lib().nann_search_eval_batch(self._h, uptr, B, ns, tk, k, C.c_void_p(ids.ctypes.data),
                                                C.c_void_p(sc.ctypes.data), C.c_void_p(nodes.ctypes.data),
                                                C.c_void_p(n.ctypes.data), C.c_void_p(status.ctypes.data), C.byref(tot),
                                                ops._stream_ptr(stream)))
        return dict(ids=ids, scores=sc, nodes=nodes, n=n, status=status, n_scored=tot.value)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.lib().nann_eval_searcher_destroy(h)
            except Exception:
                pass


class Searcher:
    """Workspace + launch sequence for batches of queries against one index and one scorer."""

    def __init__(self, index, scorer, max_batch, max_level_topn):
        self.index, self.scorer = index, scorer
        self.max_batch = int(max_batch)
        self.max_level_topn = [int(t) for t in max_level_topn]
        T = (C.c_int32 * 6)(*self.max_level_topn)
        h = C.c_void_p()
        check(_lib.lib().nann_searcher_create(index._h, scorer._h, self.max_batch, T, C.byref(h)))
        self._h = h
        self._trace = False

    def set_trace(self, enable=True):
        check(_lib.lib().nann_searcher_set_trace(self._h, int(bool(enable))))
        self._trace = bool(enable)

    def set_profile(self, enable=True):
        check(_lib.lib().nann_searcher_set_profile(self._h, int(bool(enable))))

    def profile(self):
        """-> dict(ms={score, expand_filter, topk, bitmap}, launches={...}, rows_scored, calls)"""
        ms = (C.c_double * 4)()
        ln = (C.c_int64 * 4)()
        rows, calls = C.c_int64(0), C.c_int64(0)
        check(_lib.lib().nann_searcher_get_profile(self._h, ms, ln, C.byref(rows), C.byref(calls)))
        names = ("score", "expand_filter", "topk", "bitmap")
        return dict(ms=dict(zip(names, ms[:])), launches=dict(zip(names, ln[:])), rows_scored=rows.value,
                    calls=calls.value)

    def search_device(self, users, level_topn, out_ids, out_scores, stream=None):
        """All-device form: users, out_ids (i64 [B,k]) and out_scores (f32 [B,k]) are CUDA tensors;
        nothing but the per-query status/counters crosses PCIe.  Returns (status, stats dict)."""
        uf = self.scorer.user_floats
        if not (ops._is_torch(users) and users.is_cuda):
            raise TypeError("search_device: users must be a CUDA tensor (use search() for host arrays)")
        u = users.contiguous().float().reshape(-1, uf)     # comm_seq arrives as fp16 in the reference's serving graph
        B = u.shape[0]
        T = (C.c_int32 * 6)(*[int(t) for t in level_topn])
        k = max(int(level_topn[5]), 0)
        for t, dt, name in ((out_ids, "int64", "out_ids"), (out_scores, "float32", "out_scores")):
            if not (ops._is_torch(t) and t.is_cuda and t.is_contiguous() and str(t.dtype).endswith(dt) and tuple(t.shape) == (B, k)):
                raise TypeError(f"search_device: {name} must be a contiguous CUDA {dt} tensor of shape ({B}, {k})")
        status = np.empty(B, np.int32)
        st = _lib.SearchStats()
        check(_lib.lib().nann_search_batch(self._h, C.c_void_p(u.data_ptr()), B, T, C.c_void_p(out_ids.data_ptr()),
                                           C.c_void_p(out_scores.data_ptr()), C.c_void_p(status.ctypes.data),
                                           C.byref(st), ops._stream_ptr(stream)))
        return status, dict(n_scored=np.array(st.n_scored[:], np.int64), n_expanded=np.array(st.n_expanded[:], np.int64),
                            n_failed=int(st.n_failed))

    def search_async(self, users, level_topn, out_ids, out_scores, out_status=None, stream=None):
        """Enqueue-only form: CUDA tensors in and out (out_status: optional i32 [B] CUDA tensor), no host
        synchronisation and no counters; results are complete when work ordered after it on `stream` runs."""
        uf = self.scorer.user_floats
        if not (ops._is_torch(users) and users.is_cuda and users.dtype.is_floating_point):
            raise TypeError("search_async: users must be a floating-point CUDA tensor")
        u = users.contiguous().float().reshape(-1, uf)
        B = u.shape[0]
        k = max(int(level_topn[5]), 0)
        for t, dt, shape, name in ((out_ids, "int64", (B, k), "out_ids"), (out_scores, "float32", (B, k), "out_scores"),
                                   (out_status, "int32", (B,), "out_status")):
            if t is None and name == "out_status":
                continue
            if not (ops._is_torch(t) and t.is_cuda and t.is_contiguous() and str(t.dtype).endswith(dt) and tuple(t.shape) == shape):
                raise TypeError(f"search_async: {name} must be a contiguous CUDA {dt} tensor of shape {shape}")
        T = (C.c_int32 * 6)(*[int(t) for t in level_topn])
        check(_lib.lib().nann_search_batch(self._h, C.c_void_p(u.data_ptr()), B, T, C.c_void_p(out_ids.data_ptr()),
                                           C.c_void_p(out_scores.data_ptr()),
                                           C.c_void_p(out_status.data_ptr()) if out_status is not None else None,
                                           None, ops._stream_ptr(stream)))
        self._keep = u

    def search(self, users, level_topn, stream=None):
        """users: [B, user_floats] numpy (host: H2D inside the call) or CUDA tensor.
        Returns dict(ids i64[B,k], scores f32[B,k], status i32[B], n_scored[5], n_expanded[5])."""
        uf = self.scorer.user_floats
        if ops._is_torch(users):
            u = users.contiguous().float().reshape(-1, uf)
            B, uptr = u.shape[0], C.c_void_p(u.data_ptr())
        else:
            u = np.ascontiguousarray(users, np.float32).reshape(-1, uf)
            B, uptr = u.shape[0], C.c_void_p(u.ctypes.data)
        T = (C.c_int32 * 6)(*[int(t) for t in level_topn])
        k = max(int(level_topn[5]), 0)
        ids = np.empty((B, k), np.int64)
        sc = np.empty((B, k), np.float32)
        status = np.empty(B, np.int32)
        st = _lib.SearchStats()
        check(_lib.lib().nann_search_batch(self._h, uptr, B, T, C.c_void_p(ids.ctypes.data), C.c_void_p(sc.ctypes.data),
                                           C.c_void_p(status.ctypes.data), C.byref(st), ops._stream_ptr(stream)))
        return dict(ids=ids, scores=sc, status=status, n_scored=np.array(st.n_scored[:], np.int64),
                    n_expanded=np.array(st.n_expanded[:], np.int64), n_failed=int(st.n_failed))

    def trace(self, q, rnd, cap=1 << 20):
        ids = np.empty(cap, np.int32)
        sc = np.empty(cap, np.float32)
        n = C.c_int64(0)
        check(_lib.lib().nann_searcher_get_trace(self._h, int(q), int(rnd), C.c_void_p(ids.ctypes.data),
                                                 C.c_void_p(sc.ctypes.data), cap, C.byref(n)))
        return ids[:n.value].copy(), sc[:n.value].copy()

    def nodes(self, B, k):
        out = np.empty((B, k), np.int32)
        check(_lib.lib().nann_searcher_get_nodes(self._h, C.c_void_p(out.ctypes.data), out.size))
        return out

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().nann_searcher_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
# op-by-op form: each statement is the node of build_opt_graph.py named in the comment
# ------------------------------------------------------------------------------------------------
def _fake_row_splits(x):  # build_opt_graph.py:29-30
    return np.array([0, len(x)], np.int64)


def _set_difference(a, flags):  # :33-36
    values, _rs, flags = ops.bitmap_ref_difference(a, _fake_row_splits(a), flags)
    return values, flags


def _ragged_gather(values, row_splits, idx):  # :39-49
    out, _ = ops.group_gather(values, row_splits, np.asarray(idx, np.int64), _fake_row_splits(idx), unique=False)
    return out


def _top_k(ids, scores, k):  # :52-66
    scores, indices = ops.top_k(scores, k)
    return np.asarray(ids)[indices], scores


def retrieve_opwise(scorer, user, level_topn, item_embs, item_ids, enter_points, nbr_values, nbr_row_splits):
    """build_model() (build_opt_graph.py:69-160) for one request.  item_embs may be a CUDA tensor
    (the HugeConst device copy) so the table is not re-staged on every scoring call.
    Raises NannError where session.run would fail.  Returns top_k item ids, i64 [1, level_topn[5]]."""
    T = [int(t) for t in level_topn]
    enter_points = np.asarray(enter_points, np.int32)  # :70

    def forward(idx):  # :91-107
        sc = ops.score_ids(scorer, user, item_embs, idx)
        if sc.shape[0] == 1:  # tf.squeeze -> scalar, TopKV2 then rejects rank 0 (topk_op.cc:62-65)
            raise _lib.NannError(_lib.INVALID_ARGUMENT, "input must be >= 1-D, got shape []")
        return sc

    n_items = item_ids.shape[0]
    scores = forward(enter_points)                                              # :110
    idx_results, scores_result = _top_k(enter_points, scores, T[0])             # :111
    bucket_size = int(math.ceil(n_items / 32))                                  # :115
    idx_next = _ragged_gather(nbr_values[1], nbr_row_splits[1], idx_results)    # :116
    flags = np.zeros(bucket_size, np.int32)                                     # :117-118
    idx_results, _ = _set_difference(idx_results.astype(np.int32), flags)       # :119-120
    idx_next, _ = _set_difference(idx_next, flags)                              # :121-122
    scores_next = forward(idx_next)                                             # :124
    idx_result, scores_result = _top_k(np.concatenate([idx_results, idx_next]),  # :125-127
                                       np.concatenate([scores_result, scores_next]), T[1])
    idx_candidate = idx_result.astype(np.int32)                                 # :129
    flags[:] = 0                                                                # :130-131
    idx_candidate, _ = _set_difference(idx_candidate, flags)                    # :132-133
    for i in range(3):                                                          # :135
        idx_next = _ragged_gather(nbr_values[0], nbr_row_splits[0], idx_candidate)   # :136
        idx_next, _ = _set_difference(idx_next, flags)                          # :137
        scores_next = forward(idx_next)                                         # :138
        idx_candidate, scores_candidate = _top_k(idx_next, scores_next, T[i + 2])    # :139
        idx_result = np.concatenate([idx_result, idx_candidate])                # :140
        scores_result = np.concatenate([scores_result, scores_candidate])       # :141
    idx_result, scores_result = _top_k(idx_result, scores_result, T[5])         # :143
    out_ids = ops.gather(np.asarray(item_ids, np.int64), idx_result)            # :144
    return out_ids[np.newaxis, :], scores_result                                # :149
