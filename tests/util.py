"""Shared test data: a small seeded corpus with an HNSW in the reference's file layout."""
import hashlib
import json
import os

import numpy as np

CACHE = os.environ.get("NANN_TEST_CACHE", "/tmp/nann_b200_test_cache")

SMALL = dict(n=6000, d=128, M=16, m_levels=6, seed=4, n_cand=40)
SMALL_T = [50, 100, 100, 100, 100, 100]


def build_world(n, d=128, M=16, m_levels=6, seed=4, n_cand=None, device=None, tag="w"):
    from nann_b200 import index as nix
    key = hashlib.sha1(json.dumps([n, d, M, m_levels, seed, n_cand, 3], sort_keys=True).encode()).hexdigest()[:16]
    root = os.path.join(CACHE, f"{tag}_{key}")
    embs_dir, index_dir = os.path.join(root, "embeddings"), os.path.join(root, "index")
    if not os.path.exists(os.path.join(root, "done")):
        emb = nix.synthetic_corpus(n, d, seed=0)
        item_ids = nix.synthetic_item_ids(n, seed=1)
        g = nix.build_hnsw(emb, M=M, m_levels=m_levels, seed=seed, n_cand=n_cand, device=device)
        nix.save_index(embs_dir, index_dir, emb, item_ids, g)
        open(os.path.join(root, "done"), "w").write("ok")
    emb, item_ids, g = nix.load_index_arrays(embs_dir, index_dir)
    return dict(root=root, embs_dir=embs_dir, index_dir=index_dir, emb=emb, item_ids=item_ids,
                ep=g["enter_points"], values=g["values"], row_splits=g["row_splits"])


def small_world():
    from nann_b200 import index as nix
    from nann_b200 import scorer_weights as sw
    w = build_world(tag="small", device="cpu", **SMALL)
    w["queries"] = nix.synthetic_queries(w["emb"], 64, seed=2)
    w["mlp"] = sw.mlp_weights(seed=3)
    w["T"] = list(SMALL_T)
    return w


def oracle_index(orc, w):
    return orc.Index(w["emb"], w["item_ids"], w["ep"].astype(np.int32),
                     [v.astype(np.int32) for v in w["values"]], w["row_splits"])
