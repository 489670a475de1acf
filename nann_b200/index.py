"""Index files in the reference's layout + a stand-in for the offline faiss build.

The reference builds its graph with faiss `IndexHNSWFlat(d, 32)` and dumps per-level CSR files
(NANN_impls/nann/delivery/build_hnsw_index.py:33-67).  faiss is not available here and the graph
it would produce is not pinned by any reference test, so index CONSTRUCTION is outside the parity
claim: any valid HNSW in this layout is an acceptable input and both the oracle and the CUDA path
are always fed the same files.  `build_hnsw` below produces such a graph with torch (CPU for the
small test corpora, CUDA for the 1M-row bench corpus): exact k-NN candidates per level from
blocked matmul+topk, HNSW's diversity heuristic for the forward links, reverse links, truncation
to the level's capacity (2M at level 0, M above), rows stored closest-first.  This is offline
tooling, not the hot path; it uses library matmul/topk on purpose.
"""
import math
import os

import numpy as np


# ------------------------------------------------------------------------------------------------
# synthetic data
# ------------------------------------------------------------------------------------------------
def synthetic_corpus(n, d=128, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32) / np.float32(math.sqrt(d))
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return np.ascontiguousarray(x, np.float32)


def synthetic_item_ids(n, seed=1):
    return np.random.default_rng(seed).permutation(n).astype(np.int64)  # id != row on purpose


def synthetic_queries(corpus, q, seed=2, noise=0.1):
    rng = np.random.default_rng(seed)
    rows = rng.integers(0, corpus.shape[0], q)
    g = rng.standard_normal((q, corpus.shape[1]), dtype=np.float32) / np.float32(math.sqrt(corpus.shape[1]))
    return np.ascontiguousarray(corpus[rows] + np.float32(noise) * g, np.float32)


def assign_levels(n, m_levels=32, seed=4):
    """max level per node: floor(-ln(U) / ln(m_levels)) (faiss set_default_probas)."""
    u = np.random.default_rng(seed).random(n)
    return np.floor(-np.log(np.maximum(u, 1e-300)) / math.log(m_levels)).astype(np.int32)


# ------------------------------------------------------------------------------------------------
# graph construction
# ------------------------------------------------------------------------------------------------
def _level_graph(x, nodes, cap, m_fwd, n_cand, block):
    """x: [N,d] torch (unit rows not required); nodes: LongTensor of the level's members.
    Returns (src, dst, dist) of the final links (global ids), rows closest-first."""
    import torch
    dev = x.device
    s = nodes.numel()
    if s <= 1:
        e = torch.empty(0, dtype=torch.long, device=dev)
        return e, e, torch.empty(0, device=dev)
    xs = x[nodes]
    sq = (xs * xs).sum(1)
    c = min(n_cand, s - 1)
    fsrc, fdst, fd = [], [], []
    for b0 in range(0, s, block):
        b1 = min(b0 + block, s)
        d2 = sq[b0:b1, None] + sq[None, :] - 2.0 * (xs[b0:b1] @ xs.T)
        d2[torch.arange(b1 - b0, device=dev), torch.arange(b0, b1, device=dev)] = float("inf")  # no self link
        cd, ci = torch.topk(d2, c, dim=1, largest=False, sorted=True)
        del d2
        # diversity heuristic: keep candidate j unless an already kept one is closer to it than the node is
        ce = xs[ci]                                              # [nb, c, d]
        pair = torch.cdist(ce, ce).pow(2)                        # [nb, c, c]
        keep = torch.zeros_like(ci, dtype=torch.bool)
        cnt = torch.zeros(b1 - b0, dtype=torch.long, device=dev)
        for j in range(c):
            if j == 0:
                ok = torch.ones(b1 - b0, dtype=torch.bool, device=dev)
            else:
                ok = ~((pair[:, j, :j] < cd[:, j:j + 1]) & keep[:, :j]).any(1)
            ok &= cnt < m_fwd
            keep[:, j] = ok
            cnt += ok
        rows = torch.arange(b0, b1, device=dev)[:, None].expand_as(ci)
        fsrc.append(rows[keep]); fdst.append(ci[keep]); fd.append(cd[keep])
        del ce, pair
    fsrc, fdst, fd = torch.cat(fsrc), torch.cat(fdst), torch.cat(fd)
    # forward + reverse, dedup, closest `cap` per source
    src = torch.cat([fsrc, fdst]); dst = torch.cat([fdst, fsrc]); dist = torch.cat([fd, fd])
    key = src * s + dst
    key, first = _unique_first(key)
    src, dst, dist = src[first], dst[first], dist[first]
    o = torch.argsort(dist, stable=True)
    src, dst, dist = src[o], dst[o], dist[o]
    o = torch.argsort(src, stable=True)
    src, dst, dist = src[o], dst[o], dist[o]
    counts = torch.bincount(src, minlength=s)
    starts = torch.cumsum(counts, 0) - counts
    rank = torch.arange(src.numel(), device=dev) - starts[src]
    m = rank < cap
    return nodes[src[m]], nodes[dst[m]], dist[m]


def _finish_links(nodes, s, fsrc, fdst, fd, cap):
    """forward + reverse links, dedup, closest `cap` per source (rows closest-first) -> global ids."""
    import torch
    dev = fsrc.device
    src = torch.cat([fsrc, fdst]); dst = torch.cat([fdst, fsrc]); dist = torch.cat([fd, fd])
    del fsrc, fdst, fd
    key = src * s + dst
    del src, dst
    o = torch.argsort(key, stable=True)
    key = key[o]; dist = dist[o]
    del o
    first = torch.ones_like(key, dtype=torch.bool)
    first[1:] = key[1:] != key[:-1]
    key = key[first]; dist = dist[first]
    del first
    # key is sorted by (src, dst); order each source's links by distance: sort by dist, then stably by src
    o = torch.argsort(dist, stable=True)
    key = key[o]; dist = dist[o]
    src = torch.div(key, s, rounding_mode="floor")
    o = torch.argsort(src, stable=True)
    key = key[o]; src = src[o]
    del o, dist
    counts = torch.bincount(src, minlength=s)
    starts = torch.cumsum(counts, 0) - counts
    rank = torch.arange(src.numel(), device=dev) - starts[src]
    m = rank < cap
    src = src[m]; key = key[m]
    dst = key - src * s
    return nodes[src], nodes[dst]


def _level_graph_ivf(x, nodes, cap, m_fwd, n_cand, seed=0, n_probe=8, target_cluster=2048):
    """Approximate variant of _level_graph for corpora where the exact O(s^2) candidate search is too slow
    (10M+ rows): k-means partitions, candidates = the n_cand closest points among the members of the
    n_probe nearest partitions, then the same diversity heuristic / reverse links / truncation.
    Offline tooling like _level_graph; the graph is an INPUT of the search path."""
    import torch
    dev = x.device
    s = nodes.numel()
    xs = x[nodes]
    sq = (xs * xs).sum(1)
    C = max(8, int(round(s / target_cluster)))
    g = torch.Generator(device="cpu").manual_seed(seed)
    cent = xs[torch.randperm(s, generator=g)[:C].to(dev)].clone()
    sub = xs[torch.randperm(s, generator=g)[:min(s, 64 * C)].to(dev)]

    def assign(pts, blk=262144):
        out = torch.empty(pts.shape[0], dtype=torch.long, device=dev)
        c2 = (cent * cent).sum(1)
        for b0 in range(0, pts.shape[0], blk):
            out[b0:b0 + blk] = torch.argmin(c2[None, :] - 2.0 * (pts[b0:b0 + blk] @ cent.T), dim=1)
        return out

    for _ in range(6):                                        # Lloyd iterations on a subsample
        a = assign(sub)
        cnt = torch.bincount(a, minlength=C).clamp_(min=1)
        new = torch.zeros_like(cent).index_add_(0, a, sub)
        cent = new / cnt[:, None].to(new.dtype)
    a = assign(xs)
    perm = torch.argsort(a, stable=True)                       # members of partition c: perm[off[c]:off[c+1]]
    off = torch.zeros(C + 1, dtype=torch.long, device=dev)
    off[1:] = torch.cumsum(torch.bincount(a, minlength=C), 0)
    cc = torch.cdist(cent, cent)
    probe = torch.topk(cc, min(n_probe, C), dim=1, largest=False).indices.cpu().numpy()   # includes c itself (distance 0)
    off_h = off.cpu().numpy()
    fsrc, fdst, fd = [], [], []
    for c in range(C):
        m0, m1 = int(off_h[c]), int(off_h[c + 1])
        if m1 == m0:
            continue
        mem = perm[m0:m1]
        pool = torch.cat([perm[int(off_h[p]):int(off_h[p + 1])] for p in probe[c]])
        k = min(n_cand, pool.numel() - 1)
        if k <= 0:
            continue
        d2 = sq[mem][:, None] + sq[pool][None, :] - 2.0 * (xs[mem] @ xs[pool].T)
        d2[mem[:, None] == pool[None, :]] = float("inf")      # no self link
        cd, ci = torch.topk(d2, k, dim=1, largest=False, sorted=True)
        del d2
        cg = pool[ci]                                         # level-local ids of the candidates
        ce = xs[cg]
        pair = torch.cdist(ce, ce).pow(2)
        keep = torch.zeros_like(ci, dtype=torch.bool)
        cnt = torch.zeros(mem.numel(), dtype=torch.long, device=dev)
        for j in range(k):
            if j == 0:
                ok = torch.ones(mem.numel(), dtype=torch.bool, device=dev)
            else:
                ok = ~((pair[:, j, :j] < cd[:, j:j + 1]) & keep[:, :j]).any(1)
            ok &= cnt < m_fwd
            keep[:, j] = ok
            cnt += ok
        rows = mem[:, None].expand_as(ci)
        fsrc.append(rows[keep]); fdst.append(cg[keep]); fd.append(cd[keep])
        del ce, pair
    return _finish_links(nodes, s, torch.cat(fsrc), torch.cat(fdst), torch.cat(fd), cap)


def _unique_first(key):
    import torch
    o = torch.argsort(key, stable=True)
    ks = key[o]
    first = torch.ones_like(ks, dtype=torch.bool)
    first[1:] = ks[1:] != ks[:-1]
    return ks[first], o[first]


def build_hnsw(emb, M=32, start_level=2, m_levels=None, seed=4, n_cand=None, block=2048, device=None,
               exact_limit=2_500_000):
    """-> dict(enter_points i64[n_ep], values [l] i64, row_splits [l] i64) for l < start_level,
    the arrays build_hnsw_index.py writes.  Levels with more than `exact_limit` members use the
    partitioned approximate candidate search (_level_graph_ivf)."""
    import torch
    n = emb.shape[0]
    dev = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    x = torch.as_tensor(emb, dtype=torch.float32).to(dev)
    levels = assign_levels(n, m_levels or M, seed)
    lv = torch.as_tensor(levels, device=dev)
    out = {"enter_points": np.nonzero(levels + 1 > start_level)[0].astype(np.int64), "values": [], "row_splits": [],
           "levels": levels}
    for l in range(start_level):
        nodes = torch.nonzero(lv >= l).flatten()
        cap = 2 * M if l == 0 else M
        if nodes.numel() > exact_limit:
            src, dst = _level_graph_ivf(x, nodes, cap, M, n_cand or (cap + M), seed=seed + l)
        else:
            src, dst, _ = _level_graph(x, nodes, cap, M, n_cand or (cap + M), block)
        counts = torch.bincount(src, minlength=n)
        rs = torch.zeros(n + 1, dtype=torch.long, device=dev)
        rs[1:] = torch.cumsum(counts, 0)
        out["values"].append(dst.cpu().numpy().astype(np.int64))
        out["row_splits"].append(rs.cpu().numpy().astype(np.int64))
    return out


# ------------------------------------------------------------------------------------------------
# Appendix-C files
# ------------------------------------------------------------------------------------------------
def save_index(embs_dir, index_dir, emb, item_ids, graph):
    os.makedirs(embs_dir, exist_ok=True)
    os.makedirs(index_dir, exist_ok=True)
    np.save(os.path.join(embs_dir, "item_embs.npy"), emb)
    np.save(os.path.join(embs_dir, "item_ids.npy"), np.asarray(item_ids, np.int64))
    np.save(os.path.join(index_dir, "enter_points.npy"), np.asarray(graph["enter_points"], np.int64))
    for l, (v, r) in enumerate(zip(graph["values"], graph["row_splits"])):
        np.save(os.path.join(index_dir, f"neighbors_level_{l}_values.npy"), np.asarray(v, np.int64))
        np.save(os.path.join(index_dir, f"neighbors_level_{l}_row_splits.npy"), np.asarray(r, np.int64))


def load_index_arrays(embs_dir, index_dir):
    g = {"enter_points": np.load(os.path.join(index_dir, "enter_points.npy")), "values": [], "row_splits": []}
    for l in range(2):
        g["values"].append(np.load(os.path.join(index_dir, f"neighbors_level_{l}_values.npy")))
        g["row_splits"].append(np.load(os.path.join(index_dir, f"neighbors_level_{l}_row_splits.npy")))
    emb = np.load(os.path.join(embs_dir, "item_embs.npy"))
    item_ids = np.load(os.path.join(embs_dir, "item_ids.npy"))
    return emb, item_ids, g
