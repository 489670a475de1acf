"""Distributed scoring across two real GPUs, one process per GPU (IPC-mapped windows, NVLink stores): every rank's
results must be bit-identical to the unsharded one-GPU search of the same queries.  Skipped with fewer than 2 GPUs."""
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
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import nann_b200 as nb
    from nann_b200 import builder, distributed as nd, index as nix, scorer_weights as sw
    n, B, T, n_seq = 20000, 32, [40, 60, 60, 60, 60, 40], 4
    emb = nix.synthetic_corpus(n, 128, seed=0)
    ids = nix.synthetic_item_ids(n, seed=1)
    g = builder.build_hnsw(emb, M=16, m_levels=6, seed=4, device=rank)       # the same graph on every rank
    per = -(-n // world)
    lo, hi = rank * per, min((rank + 1) * per, n)
    sc = nb.Scorer.mlp(*sw.mlp_weights(), device=rank)
    sc.set_precision(nb.SCORER_TENSOR)
    ix = nb.Index.from_arrays_sharded(n, emb[lo:hi], lo, ids, g["enter_points"], g["values"], g["row_splits"], device=rank)
    se = nb.Searcher(ix, sc, B, T)
    grp = nd.DistGroup(se, rank, world)
    grp.connect_torch()
    users = nix.synthetic_queries(emb, world * B * n_seq, seed=2)
    got_i, got_s = [], []
    for i in range(n_seq):                                                  # host outputs: blocking calls
        sc_, id_, st_ = grp.search(users[(i * world + rank) * B:(i * world + rank + 1) * B], T)
        assert np.all(st_ == 0)
        got_i.append(id_); got_s.append(sc_)
    # device outputs, back to back
    u_dev = torch.from_numpy(users).cuda()
    side = torch.cuda.Stream()
    outs = [(torch.empty((B, T[5]), dtype=torch.int64, device="cuda"), torch.empty((B, T[5]), dtype=torch.float32, device="cuda"),
             torch.empty((B,), dtype=torch.int32, device="cuda")) for _ in range(n_seq)]
    torch.cuda.synchronize()
    for rep in range(2):
        for i in range(n_seq):
            grp.search(u_dev[(i * world + rank) * B:(i * world + rank + 1) * B], T, *outs[i], stream=side)
    torch.cuda.synchronize()
    grp.check()
    for i in range(n_seq):
        np.testing.assert_array_equal(outs[i][0].cpu().numpy(), got_i[i])
        np.testing.assert_array_equal(outs[i][1].cpu().numpy().view(np.uint32), got_s[i].view(np.uint32))
    np.save(os.path.join(out_dir, f"ids_{rank}.npy"), np.stack(got_i))
    np.save(os.path.join(out_dir, f"sc_{rank}.npy"), np.stack(got_s))
    if rank == 0:                                                           # the unsharded search on one GPU
        full = nb.Index.from_arrays(emb, ids, g["enter_points"], g["values"], g["row_splits"], device=0)
        ref = nb.Searcher(full, sc, world * B * n_seq, T).search(users, T)
        np.save(os.path.join(out_dir, "ref_ids.npy"), ref["ids"])
        np.save(os.path.join(out_dir, "ref_sc.npy"), ref["scores"])
    dist.barrier()
    grp.close()
    dist.destroy_process_group()


def test_two_gpu_distributed_scoring_ipc(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world, B, n_seq, k = 2, 32, 4, 40
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ref_i = np.load(tmp_path / "ref_ids.npy").reshape(n_seq, world, B, k)
    ref_s = np.load(tmp_path / "ref_sc.npy").reshape(n_seq, world, B, k)
    for r in range(world):
        np.testing.assert_array_equal(np.load(tmp_path / f"ids_{r}.npy"), ref_i[:, r])
        np.testing.assert_array_equal(np.load(tmp_path / f"sc_{r}.npy").view(np.uint32), ref_s[:, r].view(np.uint32))
