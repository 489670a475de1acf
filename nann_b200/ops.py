"""Python mirror of the reference's op boundary (what `gen_user_ops` exposes as tf.group_gather,
tf.bitmap_ref_difference, tf.huge_const, tf.blaze_xla_op, plus stock tf.math.top_k / tf.gather):
same names, argument meaning and error behaviour, every call going through the C ABI of
libnann_b200.so (tensorflow/tensorflow/python/user_ops/user_ops.py:22-25 in the reference).

Inputs may be numpy arrays (host; staged over the GPU by the library) or CUDA torch tensors
(used in place).  Data-dependent outputs come back as numpy arrays.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import NannError, check

_NP2CODE = {np.dtype("float16"): _lib.F16, np.dtype("float32"): _lib.F32, np.dtype("float64"): _lib.F64,
            np.dtype("int32"): _lib.I32, np.dtype("int64"): _lib.I64}


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _as(x, dtype):
    """-> (pointer, n_elems, keepalive).  numpy: contiguous host array of `dtype`;
    torch: contiguous tensor, dtype must already match."""
    if x is None:
        return None, 0, None
    if _is_torch(x):
        import torch
        want = {np.dtype("int32"): torch.int32, np.dtype("int64"): torch.int64,
                np.dtype("float32"): torch.float32, np.dtype("float16"): torch.float16}[np.dtype(dtype)]
        if x.dtype != want:
            x = x.to(want)
        x = x.contiguous()
        return C.c_void_p(x.data_ptr()), x.numel(), x
    a = np.ascontiguousarray(x, dtype=dtype)
    return C.c_void_p(a.ctypes.data), a.size, a


def _stream_ptr(stream):
    if stream is None:
        return None
    return C.c_void_p(int(getattr(stream, "cuda_stream", stream)))


class _Outputs:
    """allocate_output() for the ragged ops: numpy buffers keyed by output index."""

    def __init__(self, dtypes):
        self.dtypes = dtypes
        self.arrays = {}
        self.fn = _lib.ALLOC_FN(self._alloc)

    def _alloc(self, _ctx, idx, n):
        a = np.empty(max(int(n), 0), self.dtypes[idx])
        self.arrays[idx] = a
        return a.ctypes.data if n > 0 else None


def group_gather(params_values, params_row_splits, indices_values, indices_row_splits, unique=False,
                 stream=None):
    """tf.group_gather (GroupGather_kernel.cc:18-42).  Returns (ret_values, ret_row_splits)."""
    dt = np.dtype(params_values.dtype if not _is_torch(params_values) else str(params_values.dtype).split(".")[-1])
    if dt not in (np.dtype("int32"), np.dtype("int64")):
        raise TypeError("GroupGather: T must be int32 or int64")
    pv, n_pv, k0 = _as(params_values, dt)
    prs, n_prs, k1 = _as(params_row_splits, np.int64)
    iv, n_iv, k2 = _as(indices_values, np.int64)
    irs, n_irs, k3 = _as(indices_row_splits, np.int64)
    out = _Outputs({0: dt, 1: np.int64})
    fn = _lib.lib().nann_group_gather_i32 if dt == np.int32 else _lib.lib().nann_group_gather_i64
    check(fn(pv, n_pv, prs, n_prs, iv, n_iv, irs, n_irs, int(bool(unique)), out.fn, None, _stream_ptr(stream)))
    return out.arrays.get(0, np.empty(0, dt)), out.arrays[1]


def bitmap_ref_difference(idx_next_values, idx_next_row_splits, idx_flag, stream=None):
    """tf.bitmap_ref_difference (bitmap_ops.cc:150-167).  idx_flag (int32 numpy array or CUDA
    tensor) is the Ref input: mutated in place and returned as the third output."""
    dt = np.dtype(idx_next_values.dtype if not _is_torch(idx_next_values) else str(idx_next_values.dtype).split(".")[-1])
    if dt not in (np.dtype("int32"), np.dtype("int64")):
        raise TypeError("BitmapRefDifference: T must be int32 or int64")
    if _is_torch(idx_flag):
        fl_ptr, n_fl, keep = _as(idx_flag, np.int32)
        if keep is not idx_flag and keep.data_ptr() != idx_flag.data_ptr():
            raise TypeError("idx_flag must be a contiguous int32 tensor (it is mutated in place)")
    else:
        if idx_flag.dtype != np.int32 or not idx_flag.flags["C_CONTIGUOUS"] or not idx_flag.flags["WRITEABLE"]:
            raise TypeError("idx_flag must be a writable contiguous int32 array (it is mutated in place)")
        fl_ptr, n_fl = C.c_void_p(idx_flag.ctypes.data), idx_flag.size
    v, n_v, k0 = _as(idx_next_values, dt)
    rs, n_rs, k1 = _as(idx_next_row_splits, np.int64)
    out = _Outputs({0: dt, 1: np.int64})
    fn = (_lib.lib().nann_bitmap_ref_difference_i32 if dt == np.int32
          else _lib.lib().nann_bitmap_ref_difference_i64)
    check(fn(v, n_v, rs, n_rs, fl_ptr, n_fl, out.fn, None, _stream_ptr(stream)))
    return out.arrays.get(0, np.empty(0, dt)), out.arrays[1], idx_flag


def bloom_filter_difference(idx_next_values, idx_next_row_splits, idx_flag, bucket=0, bucket_size=1, stream=None):
    """tf.bloom_filter_difference (bitmap_ops.cc:264-289): BitmapRefDifference with a 4-hash Bloom filter over idx_flag
    (int32 numpy array or CUDA tensor with at least bucket_size words; mutated in place, returned as third output)."""
    dt = np.dtype(idx_next_values.dtype if not _is_torch(idx_next_values) else str(idx_next_values.dtype).split(".")[-1])
    if dt not in (np.dtype("int32"), np.dtype("int64")):
        raise TypeError("BloomFilterDifference: T must be int32 or int64")
    if _is_torch(idx_flag):
        fl_ptr, n_fl, keep = _as(idx_flag, np.int32)
        if keep is not idx_flag and keep.data_ptr() != idx_flag.data_ptr():
            raise TypeError("idx_flag must be a contiguous int32 tensor (it is mutated in place)")
    else:
        if idx_flag.dtype != np.int32 or not idx_flag.flags["C_CONTIGUOUS"] or not idx_flag.flags["WRITEABLE"]:
            raise TypeError("idx_flag must be a writable contiguous int32 array (it is mutated in place)")
        fl_ptr, n_fl = C.c_void_p(idx_flag.ctypes.data), idx_flag.size
    v, n_v, k0 = _as(idx_next_values, dt)
    rs, n_rs, k1 = _as(idx_next_row_splits, np.int64)
    out = _Outputs({0: dt, 1: np.int64})
    fn = (_lib.lib().nann_bloom_filter_difference_i32 if dt == np.int32 else _lib.lib().nann_bloom_filter_difference_i64)
    check(fn(v, n_v, rs, n_rs, fl_ptr, n_fl, int(bucket), int(bucket_size), out.fn, None, _stream_ptr(stream)))
    return out.arrays.get(0, np.empty(0, dt)), out.arrays[1], idx_flag


def top_k(input, k, sorted=True, stream=None):  # noqa: A002 - tf.math.top_k's argument names
    """tf.math.top_k / TopKV2 on the last axis (topk_op.cc).  Returns (values f32, indices i32)."""
    if _is_torch(input):
        x = input.contiguous().float()
        shape = tuple(x.shape)
        ptr, keep = C.c_void_p(x.data_ptr()), x
    else:
        x = np.ascontiguousarray(input, np.float32)
        shape = x.shape
        ptr, keep = C.c_void_p(x.ctypes.data), x
    if len(shape) < 1:
        raise NannError(_lib.INVALID_ARGUMENT, f"input must be >= 1-D, got shape {list(shape)}")  # topk_op.cc:62-65
    k = int(k)
    cols = shape[-1]
    rows = int(np.prod(shape[:-1])) if len(shape) > 1 else 1
    kk = max(k, 0)
    values = np.empty((rows, kk), np.float32)
    indices = np.empty((rows, kk), np.int32)
    check(_lib.lib().nann_topk_v2_f32(ptr, rows, cols, k, int(bool(sorted)), C.c_void_p(values.ctypes.data),
                                      C.c_void_p(indices.ctypes.data), _stream_ptr(stream)))
    return values.reshape(shape[:-1] + (kk,)), indices.reshape(shape[:-1] + (kk,))


def batch_top_k_on_rt(values_in, row_splits_in, k, ascending=False, stream=None):
    """tf.batch_top_k_on_rt (BatchTopKOnRT_kernel.cc:25-48): ragged per-group top-k.
    k: int or per-group int64 vector.  Returns (values_out f32, idx_out i64 group-local, row_splits_out)."""
    v, n_v, k0 = _as(values_in, np.float32)
    rs, n_rs, k1 = _as(row_splits_in, np.int64)
    kk = np.atleast_1d(np.asarray(k, np.int64))
    out = _Outputs({0: np.float32, 1: np.int64, 2: np.int64})
    L = _lib.lib()
    L.nann_batch_topk_on_rt_f32.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                            C.c_int, _lib.ALLOC_FN, C.c_void_p, C.c_void_p]
    check(L.nann_batch_topk_on_rt_f32(v, n_v, rs, n_rs, C.c_void_p(kk.ctypes.data), kk.size, int(bool(ascending)),
                                      out.fn, None, _stream_ptr(stream)))
    return (out.arrays.get(0, np.empty(0, np.float32)), out.arrays.get(1, np.empty(0, np.int64)), out.arrays[2])


def _ragged2(fn32, fn64, a_vals, a_rs, b_vals, b_rs, b_dtype, stream):
    dt = np.dtype(a_vals.dtype if not _is_torch(a_vals) else str(a_vals.dtype).split(".")[-1])
    if dt not in (np.dtype("int32"), np.dtype("int64")):
        raise TypeError("T must be int32 or int64")
    av, n_av, k0 = _as(a_vals, dt)
    ars, n_ars, k1 = _as(a_rs, np.int64)
    bv, n_bv, k2 = _as(b_vals, b_dtype or dt)
    brs, n_brs, k3 = _as(b_rs, np.int64)
    out = _Outputs({0: dt, 1: np.int64})
    L = _lib.lib()
    fn = getattr(L, fn32 if dt == np.int32 else fn64)
    fn.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                   _lib.ALLOC_FN, C.c_void_p, C.c_void_p]
    check(fn(av, n_av, ars, n_ars, bv, n_bv, brs, n_brs, out.fn, None, _stream_ptr(stream)))
    return out.arrays.get(0, np.empty(0, dt)), out.arrays[1]


def batch_gather_on_rt(params_values, params_row_splits, indices_values, indices_row_splits, stream=None):
    """tf.batch_gather_on_rt (BatchGatherOnRT_kernel.cc:17-40): per group, params[group][local index]."""
    return _ragged2("nann_batch_gather_on_rt_i32", "nann_batch_gather_on_rt_i64", params_values, params_row_splits,
                    indices_values, indices_row_splits, np.int64, stream)


def batch_concat_on_rt(left_values, left_row_splits, right_values, right_row_splits, stream=None):
    """tf.batch_concat_on_rt (BatchConcatOnRT_kernel.cc:18-39): group-wise concatenation of two ragged tensors."""
    return _ragged2("nann_batch_concat_on_rt_i32", "nann_batch_concat_on_rt_i64", left_values, left_row_splits,
                    right_values, right_row_splits, None, stream)


def splits_gather(splits, indices_values, indices_row_splits, stream=None):
    """tf.splits_gather (SplitsGather_kernel.cc:20-31): expands ranges [splits[i], splits[i+1]) per group."""
    dt = np.dtype(splits.dtype if not _is_torch(splits) else str(splits.dtype).split(".")[-1])
    sp, n_sp, k0 = _as(splits, dt)
    iv, n_iv, k1 = _as(indices_values, np.int64)
    irs, n_irs, k2 = _as(indices_row_splits, np.int64)
    out = _Outputs({0: dt, 1: np.int64})
    L = _lib.lib()
    fn = L.nann_splits_gather_i32 if dt == np.int32 else L.nann_splits_gather_i64
    fn.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, _lib.ALLOC_FN, C.c_void_p, C.c_void_p]
    check(fn(sp, n_sp, iv, n_iv, irs, n_irs, out.fn, None, _stream_ptr(stream)))
    return out.arrays.get(0, np.empty(0, dt)), out.arrays[1]


def bitmap_init(idx, length, stream=None):
    """tf.bitmap_init (bitmap_ops.cc:28-43): int32 bitmap of `length` words with the bits of idx set."""
    dt = np.dtype(idx.dtype if not _is_torch(idx) else str(idx.dtype).split(".")[-1])
    v, n, k0 = _as(idx, dt)
    out = np.empty(max(int(length), 0), np.int32)
    L = _lib.lib()
    fn = L.nann_bitmap_init_i32 if dt == np.int32 else L.nann_bitmap_init_i64
    fn.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]
    check(fn(v, n, int(length), C.c_void_p(out.ctypes.data), _stream_ptr(stream)))
    return out


def bitmap_difference(idx_next, idx_flag, stream=None):
    """tf.bitmap_difference (bitmap_ops.cc:83-96): returns (idx_next_new, idx_flag_new); idx_flag is NOT mutated."""
    dt = np.dtype(idx_next.dtype if not _is_torch(idx_next) else str(idx_next.dtype).split(".")[-1])
    v, n, k0 = _as(idx_next, dt)
    fl, n_fl, k1 = _as(idx_flag, np.int32)
    new = np.empty(n_fl, np.int32)
    out = _Outputs({0: dt, 1: np.int64})
    L = _lib.lib()
    fn = L.nann_bitmap_difference_i32 if dt == np.int32 else L.nann_bitmap_difference_i64
    fn.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, _lib.ALLOC_FN, C.c_void_p, C.c_void_p]
    check(fn(v, n, fl, n_fl, C.c_void_p(new.ctypes.data), out.fn, None, _stream_ptr(stream)))
    return out.arrays.get(0, np.empty(0, dt)), new


def gather(params, indices, stream=None):
    """tf.gather(params, indices) on axis 0 (GatherV2) for a row-major table."""
    idx, n, k0 = _as(indices, np.int32)
    if _is_torch(params):
        p = params.contiguous()
        n_rows = p.shape[0]
        row_bytes = p[0].numel() * p.element_size() if p.dim() > 1 else p.element_size()
        out = np.empty((n,) + tuple(p.shape[1:]), dtype=str(p.dtype).split(".")[-1])
        ptr = C.c_void_p(p.data_ptr())
    else:
        p = np.ascontiguousarray(params)
        n_rows = p.shape[0]
        row_bytes = p[0].nbytes if p.ndim > 1 else p.itemsize
        out = np.empty((n,) + p.shape[1:], p.dtype)
        ptr = C.c_void_p(p.ctypes.data)
    check(_lib.lib().nann_gather_rows(ptr, n_rows, row_bytes, idx, n, C.c_void_p(out.ctypes.data),
                                      _stream_ptr(stream)))
    return out


class HugeConst:
    """tf.huge_const(path=, dtype=, shape=) (huge_const_op.cc:58-70): npy file -> cached tensor.
    `.numpy()` views the host copy, `.device_ptr` is the cached HBM copy (device >= 0)."""

    def __init__(self, path, dtype, shape, device=0):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(int(s) for s in shape)
        if self.dtype not in _NP2CODE:
            raise NannError(_lib.UNIMPLEMENTED, "Unsupported DataType.")
        shp = (C.c_int64 * max(len(self.shape), 1))(*self.shape)
        h = C.c_void_p()
        check(_lib.lib().nann_huge_const_create(str(path).encode(), _NP2CODE[self.dtype], shp, len(self.shape),
                                                int(device), C.byref(h)))
        self._h = h

    def numpy(self):
        n = int(np.prod(self.shape)) if self.shape else 1
        buf = (C.c_char * (n * self.dtype.itemsize)).from_address(_lib.lib().nann_huge_const_host(self._h))
        buf._owner = self  # the view keeps the handle (and so the host copy) alive
        return np.frombuffer(buf, self.dtype, n).reshape(self.shape)

    @property
    def device_ptr(self):
        return _lib.lib().nann_huge_const_device(self._h)

    @property
    def nbytes(self):
        return int(_lib.lib().nann_huge_const_bytes(self._h))

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().nann_huge_const_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def huge_const(path, dtype, shape, device=0):
    return HugeConst(path, dtype, shape, device)


def npy_peek(path):
    dt, rank = C.c_int(0), C.c_int(0)
    shp = (C.c_int64 * 8)()
    check(_lib.lib().nann_npy_peek(str(path).encode(), C.byref(dt), C.byref(rank), shp))
    code2np = {v: k for k, v in _NP2CODE.items()}
    return code2np.get(dt.value), tuple(shp[i] for i in range(rank.value))


class Scorer:
    """The model BlazeXlaOp runs.  `Scorer.mlp(...)` = synthetic 2x512 MLP (BASELINE configs 2-5);
    `Scorer.attention(blob)` = the reference's Model.forward (config 1)."""

    def __init__(self, handle, kind):
        self._h = handle
        self.kind = kind

    @classmethod
    def mlp(cls, W1, b1, W2, b2, w3, device=0):
        W1 = np.ascontiguousarray(W1, np.float32)
        H, d2 = W1.shape
        arrs = [W1] + [np.ascontiguousarray(a, np.float32) for a in (b1, W2, b2, w3)]
        h = C.c_void_p()
        check(_lib.lib().nann_scorer_create_mlp(d2 // 2, H, *[C.c_void_p(a.ctypes.data) for a in arrs],
                                                int(device), C.byref(h)))
        return cls(h, "mlp")

    @classmethod
    def attention(cls, blob, device=0):
        blob = np.ascontiguousarray(blob, np.float32)
        h = C.c_void_p()
        check(_lib.lib().nann_scorer_create_attention(C.c_void_p(blob.ctypes.data), blob.size, int(device), C.byref(h)))
        return cls(h, "attention")

    def set_admission(self, running_max=-1, max_waiting=-1, wait_ms=-1):
        """BlazeXlaOp's admission control for blaze_xla_op / score_ids (blaze_xla_kernel.cc:221-258):
        BLAZE_THREADS_NUM, DENSE_MAX_WAITING_COUNT, BlazeKernelOptions.wait_ms; -1 keeps a setting."""
        check(_lib.lib().nann_scorer_set_admission(self._h, int(running_max), int(max_waiting), int(wait_ms)))

    def admission_state(self):
        v = [C.c_int(0) for _ in range(5)]
        check(_lib.lib().nann_scorer_admission_state(self._h, *[C.byref(x) for x in v]))
        return dict(zip(("running", "waiting", "running_max", "max_waiting", "wait_ms"), [x.value for x in v]))

    def set_precision(self, precision):
        check(_lib.lib().nann_scorer_set_precision(self._h, int(precision)))

    @property
    def user_floats(self):
        return _lib.lib().nann_scorer_user_floats(self._h)

    @property
    def item_dim(self):
        return _lib.lib().nann_scorer_item_dim(self._h)

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().nann_scorer_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def blaze_xla_op(scorer, user, item_emb, stream=None, out=None):
    """tf.blaze_xla_op([user_seq_emb, item_emb], ...)[0] squeezed: logits f32[n]
    (blaze_xla_kernel.cc:24-33; build_opt_graph.py:95-107).  `out`: optional f32 [n] CUDA tensor for the logits
    (then nothing crosses PCIe and `out` is returned)."""
    u, _, k0 = _as(np.asarray(user, np.float32).reshape(-1) if not _is_torch(user) else user.float().reshape(-1), np.float32)
    if _is_torch(item_emb):
        x = item_emb.contiguous().float()
        n, ptr, k1 = x.shape[0], C.c_void_p(x.data_ptr()), x
    else:
        x = np.ascontiguousarray(item_emb, np.float32)
        n, ptr, k1 = x.shape[0], C.c_void_p(x.ctypes.data), x
    if out is not None:
        if not (_is_torch(out) and out.is_cuda and out.is_contiguous() and str(out.dtype).endswith("float32") and out.numel() == n):
            raise TypeError(f"out must be a contiguous CUDA float32 tensor with {n} elements")
        check(_lib.lib().nann_blaze_xla_run(scorer._h, u, ptr, n, C.c_void_p(out.data_ptr()), _stream_ptr(stream)))
        return out
    out = np.empty(n, np.float32)
    check(_lib.lib().nann_blaze_xla_run(scorer._h, u, ptr, n, C.c_void_p(out.ctypes.data), _stream_ptr(stream)))
    return out


def score_ids(scorer, user, table, ids, stream=None):
    """Fused tf.gather(item_embs, idx) + blaze_xla_op (build_opt_graph.py:91-107's `forward`)."""
    u, _, k0 = _as(np.asarray(user, np.float32).reshape(-1) if not _is_torch(user) else user.float().reshape(-1), np.float32)
    idx, n, k1 = _as(ids, np.int32)
    if _is_torch(table):
        t = table.contiguous()
        n_rows, ptr = t.shape[0], C.c_void_p(t.data_ptr())
    else:
        t = np.ascontiguousarray(table, np.float32)
        n_rows, ptr = t.shape[0], C.c_void_p(t.ctypes.data)
    out = np.empty(n, np.float32)
    check(_lib.lib().nann_scorer_run_ids(scorer._h, u, ptr, n_rows, idx, n, C.c_void_p(out.ctypes.data),
                                         _stream_ptr(stream)))
    return out


def merge_topk(scores, ids, k_out, stream=None, out_scores=None, out_ids=None):
    """Per-shard results [G][B][k_in] (allgather layout) -> global top k_out per query.
    With `out_scores` f32[B,k_out] / `out_ids` i64[B,k_out] CUDA tensors the result stays on the device
    (returned as given); otherwise host arrays are returned."""
    if _is_torch(scores):
        s, i = scores.contiguous().float(), ids.contiguous()
        G, B, k_in = s.shape
        sp, ip = C.c_void_p(s.data_ptr()), C.c_void_p(i.data_ptr())
    else:
        s, i = np.ascontiguousarray(scores, np.float32), np.ascontiguousarray(ids, np.int64)
        G, B, k_in = s.shape
        sp, ip = C.c_void_p(s.ctypes.data), C.c_void_p(i.ctypes.data)
    if out_scores is not None and out_ids is not None:
        assert tuple(out_scores.shape) == (B, k_out) and tuple(out_ids.shape) == (B, k_out)
        check(_lib.lib().nann_merge_topk(sp, ip, G, B, k_in, int(k_out), C.c_void_p(out_scores.data_ptr()),
                                         C.c_void_p(out_ids.data_ptr()), _stream_ptr(stream)))
        return out_scores, out_ids
    osc = np.empty((B, k_out), np.float32)
    oid = np.empty((B, k_out), np.int64)
    check(_lib.lib().nann_merge_topk(sp, ip, G, B, k_in, int(k_out), C.c_void_p(osc.ctypes.data),
                                     C.c_void_p(oid.ctypes.data), _stream_ptr(stream)))
    return osc, oid
