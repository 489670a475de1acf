"""Row-sharded retrieval over the GPUs of one box.

Each rank owns rows [lo, hi) of the corpus with their item ids and an independent HNSW; every query visits every
shard; the ONE exchange step gathers the per-shard (score f32, id i64)[B, k_s] records on every rank, followed by
the stable G-way merge (score desc, ties -> lower shard, then lower per-shard rank).

Two transports, same results:
  * `ShardGroup` (the product path): the exchange lives inside libnann_b200 -- the final top-k kernel stores its
    records straight into every rank's receive window over NVLink (CUDA IPC peer mappings) and a merge kernel on
    the group's own stream overlaps the next batch's search (nann_search_sharded, csrc/lib_shard.inl).
    torch.distributed is used once, to pass the 64-byte window handles around.
  * `sharded_search` / `allgather_results`: torch.distributed all-gather + nann_merge_topk.  Kept for transports
    without peer mappings (and for the gloo host-logic tests, which run without a GPU).
"""
import ctypes as C

import numpy as np


def shard_bounds(n, world, rank):
    """contiguous row range of `rank`; the last shards may be one row shorter or empty."""
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def shard_level_topn(T, world, scale=1.0):
    """per-shard beam widths when the corpus is split `world` ways: ceil(scale * T / world), floored so every
    TopKV2 still finds its k candidates; the final per-shard k keeps 2x slack over k/world.  `scale` is the knob
    the calibration pass turns to hold recall at the one-GPU operating point (bench.py)."""
    if world == 1:
        return [int(t) for t in T]
    t = [max(int(-(-float(x) * scale // world)), 8) for x in T[:5]]
    k = min(max(-(-int(T[5]) // world) * 2, 16), sum(t[1:5]))
    return t + [k]


def shard_beams(T, world, scales):
    """shard_level_topn with one scale per beam (T[0..4]); the calibration pass of bench.py lowers them one by one"""
    if world == 1:
        return [int(t) for t in T]
    t = [max(int(-(-float(x) * float(s) // world)), 8) for x, s in zip(T[:5], scales)]
    k = min(max(-(-int(T[5]) // world) * 2, 16), sum(t[1:5]))
    return t + [k]


def combine_status(status_by_shard):
    """[G, B] per-shard status -> [B]: the first failing shard decides (a query that fails anywhere fails)."""
    st = np.asarray(status_by_shard, np.int32)
    first = np.argmax(st != 0, axis=0)
    return st[first, np.arange(st.shape[1])]


def allgather_results(scores, ids, status=None, group=None):
    """scores f32[B,k], ids i64[B,k] torch tensors (cuda for nccl, cpu for gloo), status i32[B] optional
    -> (scores [G,B,k], ids [G,B,k], status [G,B] or None) in rank order: the layout nann_merge_topk consumes."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    B = scores.shape[0]
    # concatenated-along-dim-0 form (accepted by both gloo and nccl), viewed as [G, B, k]
    g_sc = torch.empty((world * B,) + tuple(scores.shape[1:]), dtype=scores.dtype, device=scores.device)
    g_id = torch.empty((world * B,) + tuple(ids.shape[1:]), dtype=ids.dtype, device=ids.device)
    dist.all_gather_into_tensor(g_sc, scores.contiguous(), group=group)
    dist.all_gather_into_tensor(g_id, ids.contiguous(), group=group)
    g_st = None
    if status is not None:
        st = torch.as_tensor(np.asarray(status, np.int32)).to(scores.device)
        g_st = torch.empty((world * B,), dtype=torch.int32, device=scores.device)
        dist.all_gather_into_tensor(g_st, st.contiguous(), group=group)
        g_st = g_st.view(world, B)
    return g_sc.view((world,) + tuple(scores.shape)), g_id.view((world,) + tuple(ids.shape)), g_st


def mask_failed(g_sc, g_id, g_st):
    """rows of shards that failed (0xFF fill: NaN scores, ids -1) must not take part in the merge"""
    import torch
    bad = (g_st != 0)[:, :, None]
    return (torch.where(bad, torch.full_like(g_sc, float("-inf")), g_sc),
            torch.where(bad, torch.full_like(g_id, -1), g_id))


def sharded_search(searcher, users, level_topn_shard, k_out, out_ids, out_scores, merge, group=None, stream=None):
    """One batch against this rank's shard + all-gather + merge (torch.distributed transport).
    users/out_* are CUDA tensors.  Returns (scores, ids, status [B] combined over the shards, stats): a query that
    failed on any shard fails on every rank alike (ids -1)."""
    status, stats = searcher.search_device(users, level_topn_shard, out_ids, out_scores, stream=stream)
    g_sc, g_id, g_st = allgather_results(out_scores, out_ids, status, group)
    g_sc, g_id = mask_failed(g_sc, g_id, g_st)
    m_sc, m_id = merge(g_sc, g_id, k_out)
    st = combine_status(g_st.cpu().numpy())
    if np.any(st != 0):
        bad = np.nonzero(st != 0)[0]
        if hasattr(m_id, "index_fill_"):
            import torch
            idx = torch.as_tensor(bad, device=m_id.device)
            m_id.index_fill_(0, idx, -1)
            m_sc.index_fill_(0, idx, float("nan"))
        else:
            m_id[bad] = -1
            m_sc[bad] = np.nan
    return m_sc, m_id, st, stats


class ShardGroup:
    """This rank's end of the in-library shard exchange (nann_shard_group_t)."""

    def __init__(self, rank, world, max_batch, max_k_shard, device=0):
        from . import _lib
        self._lib = _lib
        self.rank, self.world, self.device = int(rank), int(world), int(device)
        h = C.c_void_p()
        _lib.check(_lib.lib().nann_shard_group_create(self.device, self.rank, self.world, int(max_batch), int(max_k_shard),
                                                      C.byref(h)))
        self._h = h

    def export_handle(self):
        n = self._lib.lib().nann_shard_group_handle_bytes()
        buf = (C.c_ubyte * n)()
        self._lib.check(self._lib.lib().nann_shard_group_export(self._h, buf))
        return bytes(buf)

    def connect(self, handles):
        """handles: list of `world` byte strings in rank order (every rank's export_handle())."""
        blob = b"".join(handles)
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._lib.check(self._lib.lib().nann_shard_group_connect(self._h, buf))

    def connect_torch(self, group=None):
        """exchange the window handles over torch.distributed (any backend) and map the peers' windows"""
        import torch.distributed as dist
        handles = [None] * self.world
        dist.all_gather_object(handles, self.export_handle(), group=group)
        self.connect(handles)
        dist.barrier(group=group)

    @staticmethod
    def connect_local(members):
        """several members in ONE process (one per shard; same or different GPUs)"""
        from . import _lib
        arr = (C.c_void_p * len(members))(*[m._h for m in members])
        _lib.check(_lib.lib().nann_shard_group_connect_local(arr, len(members)))

    def search(self, searcher, users, level_topn_shard, k_out, out_ids=None, out_scores=None, out_status=None, stream=None):
        """users: [B, user_floats] CUDA tensor or numpy.  With CUDA output tensors (i64 [B,k_out], f32 [B,k_out],
        i32 [B]) the call only enqueues work: call wait() before reading them.  Without outputs: host arrays are
        returned (the call blocks; every rank must be in its own process or thread)."""
        from . import ops
        uf = searcher.scorer.user_floats
        if ops._is_torch(users):
            if not users.is_cuda:
                raise TypeError("users must be a CUDA tensor or a numpy array")
            u = users.contiguous().float().reshape(-1, uf)
            B, uptr = u.shape[0], C.c_void_p(u.data_ptr())
        else:
            u = np.ascontiguousarray(users, np.float32).reshape(-1, uf)
            B, uptr = u.shape[0], C.c_void_p(u.ctypes.data)
        T = (C.c_int32 * 6)(*[int(t) for t in level_topn_shard])
        k_out = int(k_out)
        host = out_ids is None
        if host:
            out_ids, out_scores = np.empty((B, k_out), np.int64), np.empty((B, k_out), np.float32)
            out_status = np.empty(B, np.int32)
            ptrs = [C.c_void_p(a.ctypes.data) for a in (out_ids, out_scores, out_status)]
        else:
            _check_out(out_ids, "int64", (B, k_out)); _check_out(out_scores, "float32", (B, k_out))
            if out_status is not None:
                _check_out(out_status, "int32", (B,))
            ptrs = [C.c_void_p(out_ids.data_ptr()), C.c_void_p(out_scores.data_ptr()),
                    C.c_void_p(out_status.data_ptr()) if out_status is not None else None]
        self._lib.check(self._lib.lib().nann_search_sharded(searcher._h, self._h, uptr, B, T, k_out, ptrs[0], ptrs[1], ptrs[2],
                                                            ops._stream_ptr(stream)))
        self._keep = u
        return out_scores, out_ids, out_status

    def push(self, searcher, users, level_topn_shard, stream=None):
        """phase 1 only (one host thread driving several members: push all of them, then merge all of them)"""
        from . import ops
        uf = searcher.scorer.user_floats
        if not (ops._is_torch(users) and users.is_cuda):
            raise TypeError("push: users must be a CUDA tensor")
        u = users.contiguous().float().reshape(-1, uf)
        T = (C.c_int32 * 6)(*[int(t) for t in level_topn_shard])
        self._lib.check(self._lib.lib().nann_search_sharded_push(searcher._h, self._h, C.c_void_p(u.data_ptr()), u.shape[0], T,
                                                                 ops._stream_ptr(stream)))
        self._keep = u

    def merge(self, k_out, out_ids, out_scores, out_status):
        """phase 2 of the pending push: CUDA output tensors, enqueue only (see wait())"""
        self._lib.check(self._lib.lib().nann_search_sharded_merge(self._h, int(k_out), C.c_void_p(out_ids.data_ptr()),
                                                                  C.c_void_p(out_scores.data_ptr()), C.c_void_p(out_status.data_ptr())))

    def wait(self, stream=None, host_block=True):
        from . import ops
        self._lib.check(self._lib.lib().nann_shard_group_wait(self._h, ops._stream_ptr(stream), int(bool(host_block))))

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.lib().nann_shard_group_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DistGroup:
    """This rank's end of DISTRIBUTED SCORING (nann_dist_group_t, csrc/lib_dist.inl): one graph, the embedding table
    row-sharded, every rank traverses its own queries and scores the candidates it owns for everybody.  Results are
    bit-identical to a one-GPU search of the unsharded index."""

    def __init__(self, searcher, rank, world):
        from . import _lib
        self._lib = _lib
        self.rank, self.world, self.searcher = int(rank), int(world), searcher
        h = C.c_void_p()
        _lib.check(_lib.lib().nann_dist_group_create(searcher._h, self.rank, self.world, C.byref(h)))
        self._h = h

    def export_handle(self):
        n = self._lib.lib().nann_shard_group_handle_bytes()
        buf = (C.c_ubyte * n)()
        self._lib.check(self._lib.lib().nann_dist_group_export(self._h, buf))
        return bytes(buf)

    def connect(self, handles):
        blob = b"".join(handles)
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        self._lib.check(self._lib.lib().nann_dist_group_connect(self._h, buf))

    def connect_torch(self, group=None):
        import torch.distributed as dist
        handles = [None] * self.world
        dist.all_gather_object(handles, self.export_handle(), group=group)
        self.connect(handles)
        dist.barrier(group=group)

    @staticmethod
    def connect_local(members):
        from . import _lib
        arr = (C.c_void_p * len(members))(*[m._h for m in members])
        _lib.check(_lib.lib().nann_dist_group_connect_local(arr, len(members)))

    def search(self, users, level_topn, out_ids=None, out_scores=None, out_status=None, stream=None, want_stats=False):
        """this rank's queries ([B, user_floats], the same B on every rank).  CUDA outputs: enqueue only (check()
        afterwards); no outputs given: host arrays are returned and the call blocks."""
        from . import ops
        se = self.searcher
        uf = se.scorer.user_floats
        if ops._is_torch(users):
            u = users.contiguous().float().reshape(-1, uf)
            B, uptr = u.shape[0], C.c_void_p(u.data_ptr())
        else:
            u = np.ascontiguousarray(users, np.float32).reshape(-1, uf)
            B, uptr = u.shape[0], C.c_void_p(u.ctypes.data)
        T = (C.c_int32 * 6)(*[int(t) for t in level_topn])
        k = max(int(level_topn[5]), 0)
        if out_ids is None:
            out_ids, out_scores, out_status = np.empty((B, k), np.int64), np.empty((B, k), np.float32), np.empty(B, np.int32)
            ptrs = [C.c_void_p(a.ctypes.data) for a in (out_ids, out_scores, out_status)]
        else:
            _check_out(out_ids, "int64", (B, k)); _check_out(out_scores, "float32", (B, k))
            if out_status is not None:
                _check_out(out_status, "int32", (B,))
            ptrs = [C.c_void_p(out_ids.data_ptr()), C.c_void_p(out_scores.data_ptr()),
                    C.c_void_p(out_status.data_ptr()) if out_status is not None else None]
        st = self._lib.SearchStats() if want_stats else None
        self._lib.check(self._lib.lib().nann_search_distributed(se._h, self._h, uptr, B, T, ptrs[0], ptrs[1], ptrs[2],
                                                                C.byref(st) if st is not None else None, ops._stream_ptr(stream)))
        self._keep = u
        if want_stats:
            return out_scores, out_ids, out_status, dict(n_scored=np.array(st.n_scored[:], np.int64), n_failed=int(st.n_failed))
        return out_scores, out_ids, out_status

    def check(self):
        self._lib.check(self._lib.lib().nann_dist_group_check(self._h))

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.lib().nann_dist_group_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _check_out(t, dtype, shape):
    if not (hasattr(t, "is_cuda") and t.is_cuda and t.is_contiguous() and str(t.dtype).endswith(dtype) and tuple(t.shape) == tuple(shape)):
        raise TypeError(f"output must be a contiguous CUDA {dtype} tensor of shape {tuple(shape)}")
