"""Row-sharded retrieval over the GPUs of one box.

Each rank owns rows [lo, hi) of the corpus with their item ids and an independent HNSW; every query
visits every shard; the ONE exchange step is an allgather of per-shard (score f32, id i64)[B, k_s],
followed by the stable G-way merge (nann_merge_topk: score desc, ties -> lower shard, then lower
per-shard rank).  Collectives go through torch.distributed (NCCL on GPUs; gloo in the CPU tests).
"""
import numpy as np


def shard_bounds(n, world, rank):
    """contiguous row range of `rank`; the last shards may be one row shorter or empty."""
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def shard_level_topn(T, world):
    """per-shard beam widths when the corpus is split `world` ways: ceil(T/world), floored so every
    TopKV2 still finds its k candidates; the final per-shard k keeps 2x slack over k/world."""
    if world == 1:
        return [int(t) for t in T]
    t = [max(-(-int(x) // world), 8) for x in T[:5]]
    k = min(max(-(-int(T[5]) // world) * 2, 16), sum(t[1:5]))
    return t + [k]


def allgather_results(scores, ids, group=None):
    """scores f32[B,k], ids i64[B,k] torch tensors (cuda for nccl, cpu for gloo)
    -> (scores [G,B,k], ids [G,B,k]) in rank order: the layout nann_merge_topk consumes."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    B = scores.shape[0]
    # concatenated-along-dim-0 form (accepted by both gloo and nccl), viewed as [G, B, k]
    g_sc = torch.empty((world * B,) + tuple(scores.shape[1:]), dtype=scores.dtype, device=scores.device)
    g_id = torch.empty((world * B,) + tuple(ids.shape[1:]), dtype=ids.dtype, device=ids.device)
    dist.all_gather_into_tensor(g_sc, scores.contiguous(), group=group)
    dist.all_gather_into_tensor(g_id, ids.contiguous(), group=group)
    return g_sc.view((world,) + tuple(scores.shape)), g_id.view((world,) + tuple(ids.shape))


def sharded_search(searcher, users, level_topn_shard, k_out, out_ids, out_scores, merge, group=None, stream=None):
    """One batch against this rank's shard + allgather + merge.  users/out_* are CUDA tensors."""
    status, stats = searcher.search_device(users, level_topn_shard, out_ids, out_scores, stream=stream)
    g_sc, g_id = allgather_results(out_scores, out_ids, group)
    m_sc, m_id = merge(g_sc, g_id, k_out)
    return m_sc, m_id, status, stats
