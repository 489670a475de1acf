"""Multi-rank host logic on CPU: world_size-2 gloo.  Each rank searches ITS shard (with the CPU
checker standing in for the GPU searcher -- this is a test of sharding, the allgather layout and the
merge rule, not of kernels), results are allgathered and merged; the merged list must equal the
top-k of the concatenation of the per-shard lists, ties -> lower shard then lower rank."""
import os
import socket

import numpy as np


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _numpy_merge(g_sc, g_id, k):
    G, B, kin = g_sc.shape
    cat_s = g_sc.transpose(1, 0, 2).reshape(B, G * kin)
    cat_i = g_id.transpose(1, 0, 2).reshape(B, G * kin)
    order = np.argsort(-cat_s, axis=1, kind="stable")[:, :k]          # position = shard*kin + rank
    return np.take_along_axis(cat_s, order, 1), np.take_along_axis(cat_i, order, 1)


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from nann_b200 import distributed as nd, index as nix, scorer_weights as sw
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, B, T = 3000, 6, [40, 60, 60, 60, 60, 40]
    full = nix.synthetic_corpus(n, 128, seed=0)
    ids = nix.synthetic_item_ids(n, seed=1)
    lo, hi = nd.shard_bounds(n, world, rank)
    emb = np.ascontiguousarray(full[lo:hi])
    g = nix.build_hnsw(emb, M=16, m_levels=4, n_cand=32, seed=4 + rank, device="cpu")
    Ts = nd.shard_level_topn(T, world)
    oix = orc.Index(emb, ids[lo:hi], g["enter_points"].astype(np.int32), [v.astype(np.int32) for v in g["values"]], g["row_splits"])
    users = nix.synthetic_queries(full, B, seed=2)                     # identical on every rank
    r = oix.search_batch_mlp(orc.Mlp(*sw.mlp_weights()), users, Ts, nthreads=1)
    assert np.all(r["status"] == 0)
    status = r["status"].copy()
    if rank == 1:
        status[2] = 3                                                   # pretend query 2 failed on shard 1 only
    g_sc, g_id, g_st = nd.allgather_results(torch.from_numpy(r["scores"]), torch.from_numpy(r["ids"]), status)
    assert tuple(g_sc.shape) == (world, B, Ts[5]) and tuple(g_st.shape) == (world, B)
    comb = nd.combine_status(g_st.numpy())                              # every rank learns of the failure
    assert comb[2] == 3 and np.all(np.delete(comb, 2) == 0)
    m_sc2, m_id2 = nd.mask_failed(g_sc, g_id, g_st)
    assert torch.all(m_id2[1, 2] == -1) and torch.all(torch.isinf(m_sc2[1, 2])) and torch.equal(m_id2[0], g_id[0])
    np.testing.assert_array_equal(g_sc[rank].numpy(), r["scores"])   # own slab sits at index `rank`
    m_sc, m_id = _numpy_merge(g_sc.numpy(), g_id.numpy(), T[5])
    np.save(os.path.join(out_dir, f"merged_{rank}.npy"), m_id)
    np.save(os.path.join(out_dir, f"local_{rank}.npy"), r["ids"])
    np.save(os.path.join(out_dir, f"localsc_{rank}.npy"), r["scores"])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_search_gloo(tmp_path):
    import torch.multiprocessing as mp
    from nann_b200 import distributed as nd
    world = 2
    assert nd.shard_bounds(10, 3, 0) == (0, 4) and nd.shard_bounds(10, 3, 2) == (8, 10) and nd.shard_bounds(2, 4, 3) == (2, 2)
    assert nd.shard_level_topn([100, 200, 200, 200, 200, 200], 1) == [100, 200, 200, 200, 200, 200]
    t8 = nd.shard_level_topn([100, 200, 200, 200, 200, 200], 8)
    assert t8[:5] == [13, 25, 25, 25, 25] and 8 * t8[5] >= 200 and t8[5] <= sum(t8[1:5])
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    m0, m1 = np.load(tmp_path / "merged_0.npy"), np.load(tmp_path / "merged_1.npy")
    np.testing.assert_array_equal(m0, m1)                              # every rank ends with the same global list
    sc = np.stack([np.load(tmp_path / f"localsc_{r}.npy") for r in range(world)])
    ids = np.stack([np.load(tmp_path / f"local_{r}.npy") for r in range(world)])
    want_s, want_i = _numpy_merge(sc, ids, 40)
    np.testing.assert_array_equal(m0, want_i)
    assert all(len(set(row.tolist())) == 40 for row in m0)             # shards are disjoint -> no duplicate ids
