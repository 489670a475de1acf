"""Multi-rank host logic on CPU: world_size-2 gloo.  Each rank searches ITS shard (with the CPU
checker standing in for the GPU searcher -- this is a test of sharding, the allgather layout and the
merge rule, not of kernels), results are allgathered and merged; the merged list must equal the
top-k of the concatenation of the per-shard lists, ties -> lower shard then lower rank.
The second test does the same for DISTRIBUTED SCORING (one graph, table row-sharded, candidates scored by their owners)."""
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


# ------------------------------------------------------------------------------------------------------------------
# distributed scoring (csrc/lib_dist.inl) on two gloo ranks: ONE graph on every rank, the embedding table row-sharded,
# every rank traverses ITS queries; per scoring round the candidates are bucketed by owner (local row index, slot =
# running count per owner), travel to the owners, are scored there from the owner's slice of the table, and the scores
# come back and are put into candidate order again.  The CPU checker stands in for the GPU kernels; what is tested is
# the data movement: ownership arithmetic, slots, the permutation, user state shared once per call.
# ------------------------------------------------------------------------------------------------------------------
def _dist_worker(rank, world, port, out_dir):
    import datetime
    import torch
    import torch.distributed as dist
    from nann_b200 import index as nix, scorer_weights as sw
    from oracle import oracle as orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world, timeout=datetime.timedelta(seconds=120))
    n, B, T, CAP = 3001, 3, [20, 40, 40, 40, 40, 40], 4096          # 3001: the last shard is shorter than the others
    full = nix.synthetic_corpus(n, 128, seed=0)
    ids = nix.synthetic_item_ids(n, seed=1)
    box = [nix.build_hnsw(full, M=16, m_levels=4, n_cand=32, seed=4, device="cpu") if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)                             # one graph, replicated
    g = box[0]
    per = -(-n // world)
    lo, hi = rank * per, min((rank + 1) * per, n)
    table_local = np.ascontiguousarray(full[lo:hi])                    # the ONLY embedding rows this rank may touch
    graph_only = orc.Index(np.zeros_like(full), ids, g["enter_points"].astype(np.int32), [v.astype(np.int32) for v in g["values"]],
                           g["row_splits"])
    om = orc.Mlp(*sw.mlp_weights())
    users = nix.synthetic_queries(full, world * B, seed=2)
    mine = users[rank * B:(rank + 1) * B]

    def gather(x):
        out = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(out, x)
        return [o.numpy() for o in out]

    got_ids, got_sc, rows_served = [], [], 0
    for q in range(B):
        u_all = gather(torch.from_numpy(mine[q].copy()))               # the users' state, once per call (dist_bcast_kernel)

        def score_round(rnd, cand):
            nonlocal rows_served
            cand = np.asarray(cand, np.int64)
            assert len(cand) <= CAP
            req = np.full((world, CAP), -1, np.int32)                  # req[owner][slot] = LOCAL row index
            cnt = np.zeros(world, np.int32)
            perm = np.zeros((len(cand), 2), np.int64)
            for i, c in enumerate(cand):                               # dist_bucket_push_kernel (slot order is arbitrary there)
                o = int(c // per)
                perm[i] = (o, cnt[o])
                req[o, cnt[o]] = c - o * per
                cnt[o] += 1
            all_req = gather(torch.from_numpy(req))                    # ids out            [src][owner][slot]
            all_cnt = gather(torch.from_numpy(cnt))
            resp = np.zeros((world, CAP), np.float32)                  # resp[src][slot]: what I own, for every source
            for src in range(world):
                k = int(all_cnt[src][rank])
                rows = all_req[src][rank][:k]
                assert np.all((rows >= 0) & (rows < hi - lo))
                if k:
                    resp[src, :k] = om.score(u_all[src], table_local, rows.astype(np.int32))
                rows_served += k
            all_resp = gather(torch.from_numpy(resp))                  # scores back        [owner][src][slot]
            return np.array([all_resp[o][rank][j] for o, j in perm], np.float32)   # dist_unbucket_kernel

        r = graph_only.search(score_round, T)
        assert r["status"] == 0
        got_ids.append(r["ids"]); got_sc.append(r["scores"])
    np.save(os.path.join(out_dir, f"dist_ids_{rank}.npy"), np.stack(got_ids))
    np.save(os.path.join(out_dir, f"dist_sc_{rank}.npy"), np.stack(got_sc))
    np.save(os.path.join(out_dir, f"dist_served_{rank}.npy"), np.array([rows_served]))
    if rank == 0:                                                       # the unsharded search, for the comparison
        whole = orc.Index(full, ids, g["enter_points"].astype(np.int32), [v.astype(np.int32) for v in g["values"]], g["row_splits"])
        w = whole.search_batch_mlp(om, users, T, nthreads=1)
        assert np.all(w["status"] == 0)
        np.save(os.path.join(out_dir, "dist_want_ids.npy"), w["ids"])
        np.save(os.path.join(out_dir, "dist_want_sc.npy"), w["scores"])
        np.save(os.path.join(out_dir, "dist_want_rows.npy"), np.array([w["n_scored"]]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_distributed_scoring_gloo(tmp_path):
    """ids and scores of every rank's queries equal the unsharded search bit for bit, although no rank holds more than its
    slice of the table; the two ranks together scored exactly the rows of the unsharded search (work is split, not repeated)"""
    import torch.multiprocessing as mp
    world, B = 2, 3
    mp.spawn(_dist_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    want_i, want_s = np.load(tmp_path / "dist_want_ids.npy"), np.load(tmp_path / "dist_want_sc.npy")
    for r in range(world):
        np.testing.assert_array_equal(np.load(tmp_path / f"dist_ids_{r}.npy"), want_i[r * B:(r + 1) * B])
        np.testing.assert_array_equal(np.load(tmp_path / f"dist_sc_{r}.npy").view(np.uint32), want_s[r * B:(r + 1) * B].view(np.uint32))
    served = sum(int(np.load(tmp_path / f"dist_served_{r}.npy")[0]) for r in range(world))
    assert served == int(np.load(tmp_path / "dist_want_rows.npy")[0])
    assert min(int(np.load(tmp_path / f"dist_served_{r}.npy")[0]) for r in range(world)) > 0.3 * served   # both ranks carry load
