"""ctypes front-end of the CPU oracle (oracle/nann_oracle.c).

TEST INFRASTRUCTURE ONLY.  May be imported from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never from nann_b200/ (the product).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnann_oracle.so")

OK, INVALID_ARGUMENT, NOT_FOUND, UNIMPLEMENTED, INTERNAL = 0, 3, 5, 12, 13
DTYPE_CODE = {np.dtype("float16"): 0, np.dtype("float32"): 1, np.dtype("float64"): 2,
              np.dtype("int32"): 3, np.dtype("int64"): 4}


class OracleError(Exception):
    def __init__(self, code, msg=""):
        super().__init__(f"oracle status {code} {msg}")
        self.code = code


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("nann_oracle.c", "nann_oracle.h", "Makefile")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src)):
        return _SO
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_version.restype = C.c_char_p
        _lib.orc_mlp_create.restype = C.c_void_p
        _lib.orc_attn_create.restype = C.c_void_p
        _lib.orc_attn_blob_size.restype = C.c_int64
        _lib.orc_search_batch_mlp.restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def validate_ragged(n_values, row_splits):
    rs = _c(row_splits, np.int64)
    return lib().orc_validate_ragged(C.c_int64(n_values), _p(rs), C.c_int64(rs.size))


def group_gather(params_values, params_row_splits, indices_values, indices_row_splits, unique=False):
    """Mirrors tf.group_gather (GroupGather op).  Returns (ret_values, ret_row_splits)."""
    pv = np.ascontiguousarray(params_values)
    assert pv.dtype in (np.int32, np.int64)
    prs, iv, irs = _c(params_row_splits, np.int64), _c(indices_values, np.int64), _c(indices_row_splits, np.int64)
    fn = lib().orc_group_gather_i32 if pv.dtype == np.int32 else lib().orc_group_gather_i64
    n_ret, n_rrs, code = C.c_int64(0), C.c_int64(0), C.c_int(0)
    rrs = np.zeros(max(irs.size, 1), np.int64)
    args = [_p(pv), C.c_int64(pv.size), _p(prs), C.c_int64(prs.size), _p(iv), C.c_int64(iv.size),
            _p(irs), C.c_int64(irs.size), C.c_int(int(unique))]
    st = fn(*args, None, C.byref(n_ret), _p(rrs), C.byref(n_rrs), C.byref(code))
    if st != OK:
        raise OracleError(st, f"Invalid RaggedTensor, code: {code.value}")
    out = np.zeros(n_ret.value, pv.dtype)
    st = fn(*args, _p(out), C.byref(n_ret), _p(rrs), C.byref(n_rrs), C.byref(code))
    if st != OK:
        raise OracleError(st)
    return out[:n_ret.value], rrs[:n_rrs.value].copy()


def bitmap_ref_difference(values, row_splits, flags, check_bounds=True):
    """Mirrors tf.bitmap_ref_difference.  flags (int32 ndarray) is mutated in place."""
    v = np.ascontiguousarray(values)
    assert v.dtype in (np.int32, np.int64)
    assert flags.dtype == np.int32 and flags.flags["C_CONTIGUOUS"]
    rs = _c(row_splits, np.int64)
    fn = lib().orc_bitmap_ref_difference_i32 if v.dtype == np.int32 else lib().orc_bitmap_ref_difference_i64
    out = np.zeros(max(v.size, 1), v.dtype)
    crs = np.zeros(max(rs.size, 1), np.int64)
    n_c, n_crs, code = C.c_int64(0), C.c_int64(0), C.c_int(0)
    st = fn(_p(v), C.c_int64(v.size), _p(rs), C.c_int64(rs.size), _p(flags), C.c_int64(flags.size),
            C.c_int(int(check_bounds)), _p(out), C.byref(n_c), _p(crs), C.byref(n_crs), C.byref(code))
    if st != OK:
        raise OracleError(st, f"code: {code.value}")
    return out[:n_c.value].copy(), crs[:n_crs.value].copy(), flags


def top_k(inp, k):
    """Mirrors tf.math.top_k(input, k) (TopKV2, sorted=True) on the last axis."""
    x = _c(inp, np.float32)
    if x.ndim < 1:
        raise OracleError(INVALID_ARGUMENT, "input must be >= 1-D")
    cols = x.shape[-1]
    rows = int(np.prod(x.shape[:-1])) if x.ndim > 1 else 1
    vals = np.zeros((rows, max(k, 0)), np.float32)
    idx = np.zeros((rows, max(k, 0)), np.int32)
    st = lib().orc_topk_v2_f32(_p(x), C.c_int64(rows), C.c_int64(cols), C.c_int(k), _p(vals), _p(idx))
    if st != OK:
        raise OracleError(st, "Need k >= 0" if k < 0 else "input must have at least k columns")
    shp = x.shape[:-1] + (k,)
    return vals.reshape(shp), idx.reshape(shp)


def batch_top_k_on_rt(values, row_splits, k, ascending=False):
    """Mirrors tf.batch_top_k_on_rt.  Returns (values_out, idx_out (group-local), row_splits_out)."""
    v, rs = _c(values, np.float32), _c(row_splits, np.int64)
    kk = np.atleast_1d(np.asarray(k, np.int64))
    vo, io = np.zeros(max(v.size, 1), np.float32), np.zeros(max(v.size, 1), np.int64)
    ro = np.zeros(max(rs.size, 1), np.int64)
    n, code = C.c_int64(0), C.c_int(0)
    st = lib().orc_batch_topk_on_rt_f32(_p(v), C.c_int64(v.size), _p(rs), C.c_int64(rs.size), _p(kk), C.c_int64(kk.size),
                                        C.c_int(int(ascending)), _p(vo), _p(io), _p(ro), C.byref(n), C.byref(code))
    if st != OK:
        raise OracleError(st, f"code: {code.value}")
    return vo[:n.value].copy(), io[:n.value].copy(), ro[:max(rs.size, 1)].copy()


def huge_const_check(path, dtype, shape, read=False):
    dt = np.dtype(dtype)
    shp = np.asarray(shape, np.int64)
    dst = np.zeros(int(np.prod(shape)), dt) if read else None
    st = lib().orc_huge_const_load(path.encode(), C.c_int(DTYPE_CODE[dt]), _p(shp), C.c_int(shp.size),
                                   _p(dst), C.c_int64(dst.nbytes if dst is not None else 0))
    return st, (dst.reshape(shape) if dst is not None else None)


class Mlp:
    """mlp2x512-style scorer (fp32 definition: sequential fmaf chains)."""

    def __init__(self, W1, b1, W2, b2, w3):
        self.W1, self.b1, self.W2 = _c(W1, np.float32), _c(b1, np.float32), _c(W2, np.float32)
        self.b2, self.w3 = _c(b2, np.float32), _c(w3, np.float32)
        self.H = self.W1.shape[0]
        self.d = self.W1.shape[1] // 2
        assert self.W2.shape == (self.H, self.H)
        self.h = lib().orc_mlp_create(C.c_int(self.d), C.c_int(self.H), _p(self.W1), _p(self.b1),
                                      _p(self.W2), _p(self.b2), _p(self.w3))
        assert self.h

    def __del__(self):
        if getattr(self, "h", None):
            try:                      # module globals may already be gone at interpreter shutdown
                lib().orc_mlp_destroy(C.c_void_p(self.h))
            except Exception:
                pass
            self.h = None

    def score_def(self, u, x):
        u, x = _c(u, np.float32), _c(x, np.float32)
        out = np.zeros(x.shape[0], np.float32)
        lib().orc_mlp_score_def(C.c_void_p(self.h), _p(u), _p(x), C.c_int64(x.shape[0]), _p(out))
        return out

    def score(self, u, table, ids=None):
        u, table = _c(u, np.float32), _c(table, np.float32)
        if ids is not None:
            ids = _c(ids, np.int32)
            n = ids.size
        else:
            n = table.shape[0]
        out = np.zeros(n, np.float32)
        lib().orc_mlp_score(C.c_void_p(self.h), _p(u), _p(table), _p(ids), C.c_int64(n), _p(out))
        return out


class Attn:
    """Reference scorer (config 1): attention + 4-layer DNN, fp32."""
    L, E = 50, 64

    def __init__(self, blob):
        self.blob = _c(blob, np.float32)
        assert self.blob.size == lib().orc_attn_blob_size()
        self.h = lib().orc_attn_create(_p(self.blob), C.c_int64(self.blob.size))
        assert self.h

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_attn_destroy(C.c_void_p(self.h))
            self.h = None

    def score(self, user, table, ids=None):
        user, table = _c(user, np.float32).reshape(50, 64), _c(table, np.float32)
        if ids is not None:
            ids = _c(ids, np.int32)
            n = ids.size
        else:
            n = table.shape[0]
        out = np.zeros(n, np.float32)
        lib().orc_attn_score(C.c_void_p(self.h), _p(user), _p(table), _p(ids), C.c_int64(n), _p(out))
        return out


class _IndexStruct(C.Structure):
    _fields_ = [("n_items", C.c_int64), ("dim", C.c_int), ("emb", C.c_void_p),
                ("item_ids", C.c_void_p), ("ep", C.c_void_p), ("n_ep", C.c_int64),
                ("nbr_values", C.c_void_p * 2), ("nbr_row_splits", C.c_void_p * 2)]


class _Stats(C.Structure):
    _fields_ = [("n_scored", C.c_int64 * 5), ("n_expanded", C.c_int64 * 5)]


_SCORE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.c_int64, C.POINTER(C.c_float))


class Index:
    """Appendix-C layout held as numpy arrays (emb f32[N,d], item_ids i64[N], ep i32, CSR l0/l1)."""

    def __init__(self, emb, item_ids, enter_points, nbr_values, nbr_row_splits):
        self.emb = _c(emb, np.float32)
        self.item_ids = _c(item_ids, np.int64)
        self.ep = _c(enter_points, np.int32)
        self.nv = [_c(v, np.int32) for v in nbr_values]
        self.nrs = [_c(r, np.int64) for r in nbr_row_splits]
        s = _IndexStruct()
        s.n_items, s.dim = self.emb.shape[0], self.emb.shape[1]
        s.emb, s.item_ids, s.ep, s.n_ep = self.emb.ctypes.data, self.item_ids.ctypes.data, self.ep.ctypes.data, self.ep.size
        for l in range(2):
            s.nbr_values[l] = self.nv[l].ctypes.data
            s.nbr_row_splits[l] = self.nrs[l].ctypes.data
        self.c = s

    def search(self, score_fn, level_topn, trace=False, trace_cap=1 << 20):
        """exec.pb dataflow for one query.  score_fn(round, ids ndarray) -> f32 scores.
        Returns dict(status, ids, scores, nodes, n_scored, n_expanded[, trace])."""
        T = _c(level_topn, np.int32)
        k = int(T[5])
        out_ids, out_sc, out_nodes = np.zeros(k, np.int64), np.zeros(k, np.float32), np.zeros(k, np.int32)

        def cb(_ctx, rnd, ids_p, n, out_p):
            ids = np.ctypeslib.as_array(ids_p, shape=(n,)) if n > 0 else np.zeros(0, np.int32)
            sc = np.asarray(score_fn(rnd, ids.copy()), np.float32)
            assert sc.shape == (n,), (sc.shape, n)
            if n > 0:
                np.ctypeslib.as_array(out_p, shape=(n,))[:] = sc

        cbf = _SCORE_FN(cb)
        st = _Stats()
        if trace:
            tids = [np.zeros(trace_cap, np.int32) for _ in range(5)]
            tsc = [np.zeros(trace_cap, np.float32) for _ in range(5)]
            tn = np.zeros(5, np.int64)
            tidp = (C.c_void_p * 5)(*[a.ctypes.data for a in tids])
            tscp = (C.c_void_p * 5)(*[a.ctypes.data for a in tsc])
            rc = lib().orc_search(C.byref(self.c), cbf, None, _p(T), _p(out_ids), _p(out_sc), _p(out_nodes),
                                  C.byref(st), tidp, tscp, _p(tn), C.c_int64(trace_cap))
        else:
            rc = lib().orc_search(C.byref(self.c), cbf, None, _p(T), _p(out_ids), _p(out_sc), _p(out_nodes),
                                  C.byref(st), None, None, None, C.c_int64(0))
        res = dict(status=rc, ids=out_ids, scores=out_sc, nodes=out_nodes,
                   n_scored=np.array(st.n_scored[:], np.int64), n_expanded=np.array(st.n_expanded[:], np.int64))
        if trace:
            res["trace"] = [(tids[r][:tn[r]].copy(), tsc[r][:tn[r]].copy()) for r in range(5)]
        return res

    def search_eval(self, score_fn, num_scoring_per_level=(3, 1, 1), top_k_per_level=(400, 200, 100), topk_eval=200):
        """`main.py --job-type test` traversal (model.py:299-362) for one query.
        score_fn(round, ids ndarray) -> f32 scores.  Returns dict(status, n, ids, scores, nodes, n_scored)."""
        ns, tk = _c(num_scoring_per_level, np.int32), _c(top_k_per_level, np.int32)
        k = int(topk_eval)
        out_ids, out_sc, out_nodes = np.full(k, -1, np.int64), np.zeros(k, np.float32), np.full(k, -1, np.int32)

        def cb(_ctx, rnd, ids_p, n, out_p):
            ids = np.ctypeslib.as_array(ids_p, shape=(n,)) if n > 0 else np.zeros(0, np.int32)
            sc = np.asarray(score_fn(rnd, ids.copy()), np.float32)
            assert sc.shape == (n,), (sc.shape, n)
            if n > 0:
                np.ctypeslib.as_array(out_p, shape=(n,))[:] = sc

        cbf = _SCORE_FN(cb)
        n_out, tot = C.c_int32(0), C.c_int64(0)
        rc = lib().orc_search_eval(C.byref(self.c), cbf, None, _p(ns), _p(tk), C.c_int(k), _p(out_ids), _p(out_sc),
                                   _p(out_nodes), C.byref(n_out), C.byref(tot))
        return dict(status=rc, n=n_out.value, ids=out_ids, scores=out_sc, nodes=out_nodes, n_scored=tot.value)

    def search_batch_mlp(self, mlp, users, level_topn, nthreads=0):
        """Request-parallel batch (one request per core).  Returns dict(ids, scores, status, seconds, n_scored)."""
        users = _c(users, np.float32)
        B = users.shape[0]
        T = _c(level_topn, np.int32)
        k = int(T[5])
        ids, sc, status = np.zeros((B, k), np.int64), np.zeros((B, k), np.float32), np.zeros(B, np.int32)
        tot = C.c_int64(0)
        secs = lib().orc_search_batch_mlp(C.byref(self.c), C.c_void_p(mlp.h), _p(users), C.c_int64(B), _p(T),
                                          C.c_int(nthreads), _p(ids), _p(sc), _p(status), C.byref(tot))
        return dict(ids=ids, scores=sc, status=status, seconds=secs, n_scored=tot.value)


def build_hnsw(emb, levels, M=32, n_levels=2, nthreads=0):
    """CPU statement of the CUDA index builder (nann_hnsw_build): -> (values [l] i32, row_splits [l] i64 [n+1]).
    levels[i] = highest level of node i (0-based)."""
    emb = _c(emb, np.float32)
    n, d = emb.shape
    levels = np.asarray(levels)
    nthreads = nthreads or min(os.cpu_count() or 1, 32)
    values, row_splits = [], []
    for l in range(n_levels):
        nodes = np.nonzero(levels >= l)[0].astype(np.int64)
        s = len(nodes)
        cap = 2 * M if l == 0 else M
        counts = np.zeros(n, np.int64)
        rows = {}
        if s >= 2:
            X = _c(emb[nodes], np.float32)
            links = np.zeros((s, cap), np.int32)
            cnt = np.zeros(s, np.int32)
            st = lib().orc_build_level(_p(X), C.c_int64(s), int(d), int(min(cap + M, s - 1)), int(M), int(cap), int(nthreads),
                                       _p(links), _p(cnt))
            if st != OK:
                raise OracleError(st, "orc_build_level")
            counts[nodes] = cnt
            vals = np.concatenate([nodes[links[i, :cnt[i]]] for i in range(s)]) if cnt.sum() else np.zeros(0, np.int64)
        else:
            vals = np.zeros(0, np.int64)
        rs = np.zeros(n + 1, np.int64)
        rs[1:] = np.cumsum(counts)
        values.append(vals.astype(np.int32))
        row_splits.append(rs)
    return values, row_splits


def fingerprint64(b):
    """farmhash::Fingerprint64 (inputs up to 32 bytes)"""
    lib().orc_fingerprint64.restype = C.c_uint64
    b = bytes(b)
    return int(lib().orc_fingerprint64(C.c_char_p(b), C.c_int64(len(b))))


def bloom_filter_difference(idx_next_values, idx_next_row_splits, idx_flag, bucket=0, bucket_size=1):
    """Mirrors tf.bloom_filter_difference (BloomFilterDifference op); idx_flag (int32 array) is mutated in place.
    Returns (c_values, c_row_splits, idx_flag)."""
    v = np.ascontiguousarray(idx_next_values)
    assert v.dtype in (np.int32, np.int64)
    rs = _c(idx_next_row_splits, np.int64)
    assert idx_flag.dtype == np.int32 and idx_flag.flags["C_CONTIGUOUS"]
    out = np.empty(max(v.size, 1), v.dtype)
    ors = np.empty(max(rs.size, 1), np.int64)
    n_c, code = C.c_int64(0), C.c_int(0)
    fn = lib().orc_bloom_filter_difference_i32 if v.dtype == np.int32 else lib().orc_bloom_filter_difference_i64
    st = fn(_p(v), C.c_int64(v.size), _p(rs), C.c_int64(rs.size), _p(idx_flag), C.c_int64(idx_flag.size), C.c_int64(bucket),
            C.c_int64(bucket_size), _p(out), _p(ors), C.byref(n_c), C.byref(code))
    if st != OK:
        raise OracleError(st, f"Invalid RaggedTensor input0 a, code: {code.value}" if code.value else "bad bloom arguments")
    return out[:n_c.value].copy(), ors[:rs.size].copy(), idx_flag
