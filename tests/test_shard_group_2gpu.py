"""Sharded search across two real GPUs, one process per GPU: window handles over torch.distributed, peer windows
mapped with CUDA IPC, records pushed over NVLink by the final top-k kernel, merge on every rank.  The merged
result must equal the stable merge of the per-shard ORACLE results.  Skipped on boxes with fewer than 2 GPUs
(run it with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)      # plumbing only: 64-byte handles + barriers
    import nann_b200 as nb
    from nann_b200 import distributed as nd, index as nix, scorer_weights as sw
    from oracle import oracle as orc
    n, B, T, n_seq = 6000, 16, [40, 60, 60, 60, 60, 40], 5
    full = nix.synthetic_corpus(n, 128, seed=0)
    ids = nix.synthetic_item_ids(n, seed=1)
    lo, hi = nd.shard_bounds(n, world, rank)
    emb = np.ascontiguousarray(full[lo:hi])
    g = nix.build_hnsw(emb, M=16, m_levels=4, n_cand=32, seed=4 + rank, device="cpu")
    Ts = nd.shard_level_topn(T, world)
    W = sw.mlp_weights()
    ix = nb.Index.from_arrays(emb, ids[lo:hi], g["enter_points"], g["values"], g["row_splits"], device=rank)
    sc = nb.Scorer.mlp(*W, device=rank)
    se = nb.Searcher(ix, sc, B, Ts)
    grp = nd.ShardGroup(rank, world, B, Ts[5], device=rank)
    grp.connect_torch()
    users = nix.synthetic_queries(full, B * n_seq, seed=2)             # identical on every rank
    oix = orc.Index(emb, ids[lo:hi], g["enter_points"].astype(np.int32), [v.astype(np.int32) for v in g["values"]], g["row_splits"])
    local = oix.search_batch_mlp(orc.Mlp(*W), users, Ts, nthreads=4)
    assert np.all(local["status"] == 0)
    np.save(os.path.join(out_dir, f"local_sc_{rank}.npy"), local["scores"])
    np.save(os.path.join(out_dir, f"local_id_{rank}.npy"), local["ids"])
    # (a) host outputs: one blocking call per batch
    host = [grp.search(se, users[i * B:(i + 1) * B], Ts, T[5]) for i in range(n_seq)]
    np.save(os.path.join(out_dir, f"host_sc_{rank}.npy"), np.stack([h[0] for h in host]))
    np.save(os.path.join(out_dir, f"host_id_{rank}.npy"), np.stack([h[1] for h in host]))
    assert all(np.all(h[2] == 0) for h in host)
    # (b) device outputs: everything enqueued back to back, exchange + merge overlap the next batch's search
    u_dev = torch.from_numpy(users).cuda()
    outs = [(torch.empty((B, T[5]), dtype=torch.int64, device="cuda"), torch.empty((B, T[5]), dtype=torch.float32, device="cuda"),
             torch.empty((B,), dtype=torch.int32, device="cuda")) for _ in range(n_seq)]
    side = torch.cuda.Stream()
    torch.cuda.synchronize()
    for rep in range(3):
        for i in range(n_seq):
            grp.search(se, u_dev[i * B:(i + 1) * B], Ts, T[5], *outs[i], stream=side)
        grp.wait()
    torch.cuda.synchronize()
    np.save(os.path.join(out_dir, f"dev_id_{rank}.npy"), np.stack([o[0].cpu().numpy() for o in outs]))
    np.save(os.path.join(out_dir, f"dev_sc_{rank}.npy"), np.stack([o[1].cpu().numpy() for o in outs]))
    dist.barrier()
    grp.close()
    dist.destroy_process_group()


def test_two_gpu_sharded_search_ipc(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world, B, n_seq, k = 2, 16, 5, 40
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sc = np.stack([np.load(tmp_path / f"local_sc_{r}.npy") for r in range(world)])     # [G, B*n_seq, k_s]
    ids = np.stack([np.load(tmp_path / f"local_id_{r}.npy") for r in range(world)])
    G, Q, kin = sc.shape
    cat_s = sc.transpose(1, 0, 2).reshape(Q, G * kin)
    cat_i = ids.transpose(1, 0, 2).reshape(Q, G * kin)
    order = np.argsort(-cat_s, axis=1, kind="stable")[:, :k]
    want_s, want_i = np.take_along_axis(cat_s, order, 1), np.take_along_axis(cat_i, order, 1)
    for r in range(world):
        for kind in ("host", "dev"):
            got_i = np.load(tmp_path / f"{kind}_id_{r}.npy").reshape(Q, k)
            got_s = np.load(tmp_path / f"{kind}_sc_{r}.npy").reshape(Q, k)
            np.testing.assert_array_equal(got_i, want_i, err_msg=f"rank {r} {kind}")
            np.testing.assert_array_equal(got_s.view(np.uint32), want_s.view(np.uint32))
