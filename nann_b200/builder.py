"""HNSW index construction on the GPU (nann_hnsw_build, csrc/builder_kernels.cuh): the files
NANN_impls/nann/delivery/build_hnsw_index.py:33-67 dumps, in well under a second per million rows.

Same return value as nann_b200.index.build_hnsw (the torch stand-in it replaces): enter_points i64, per level
`values` i64 and `row_splits` i64 [n+1].  Levels are drawn with faiss's distribution (index.assign_levels)."""
import ctypes as C

import numpy as np

from . import _lib, index as _index
from ._lib import check


class BuildStats(C.Structure):
    _fields_ = [("seconds_knn", C.c_double), ("seconds_links", C.c_double), ("n_forward_links", C.c_int64),
                ("n_overflow", C.c_int64)]


def build_hnsw(emb, M=32, start_level=2, m_levels=None, seed=4, device=0, levels=None, return_stats=False, values_dtype=np.int64):
    """emb: float32 [n, 128] numpy array or CUDA tensor.  m_levels: base of the level distribution (default M).
    values_dtype: int64 like the files build_hnsw_index.py writes, or int32 (what build_opt_graph.py casts them to;
    half the host memory for a 12.5M-row shard)."""
    is_torch = type(emb).__module__.startswith("torch")
    if is_torch:
        e = emb.contiguous().float()
        n, d = e.shape
        ptr = C.c_void_p(e.data_ptr())
    else:
        e = np.ascontiguousarray(emb, np.float32)
        n, d = e.shape
        ptr = C.c_void_p(e.ctypes.data)
    if levels is None:
        levels = _index.assign_levels(n, m_levels or M, seed)
    lv = np.ascontiguousarray(levels, np.int32)
    outs = {}

    def _alloc(_ctx, idx, cnt):
        a = np.empty(max(int(cnt), 0), np.int32 if idx % 2 == 0 else np.int64)
        outs[idx] = a
        return a.ctypes.data if a.size else 0

    cb = _lib.ALLOC_FN(_alloc)
    st = BuildStats()
    L = _lib.lib()
    L.nann_hnsw_build.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, _lib.ALLOC_FN,
                                  C.c_void_p, C.c_void_p]
    check(L.nann_hnsw_build(ptr, n, d, C.c_void_p(lv.ctypes.data), int(M), int(start_level), int(device), cb, None, C.byref(st)))
    out = {"enter_points": np.nonzero(lv + 1 > start_level)[0].astype(np.int64), "levels": lv,
           "values": [outs[2 * l].astype(values_dtype, copy=False) for l in range(start_level)],
           "row_splits": [outs[2 * l + 1] for l in range(start_level)]}
    if return_stats:
        out["stats"] = dict(seconds_knn=st.seconds_knn, seconds_links=st.seconds_links, n_forward_links=st.n_forward_links,
                            n_overflow=st.n_overflow)
    return out
