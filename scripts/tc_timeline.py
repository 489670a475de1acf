"""Per-tile pipeline timeline of the tensor-core scorer (CTA 0), from clock64() stamps."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import nann_b200 as nb
from nann_b200 import _lib, index as nix, scorer_weights as sw

emb = nix.synthetic_corpus(400000, 128, seed=0)
emb_d = torch.from_numpy(emb).cuda()
sc = nb.Scorer.mlp(*sw.mlp_weights()); sc.set_precision(nb.SCORER_TENSOR)
rng = np.random.default_rng(0)
n = 148 * 128 * 40                        # 40 tiles per CTA
ids = torch.from_numpy(rng.integers(0, emb.shape[0], n).astype(np.int32)).cuda()
u = nix.synthetic_queries(emb, 1, seed=1)[0]
nb.score_ids(sc, u, emb_d, ids)           # warm
buf = torch.zeros(64 * 48, dtype=torch.int64, device="cuda")
L = _lib.lib(); L.nann_debug_tc_trace.argtypes = [C.c_void_p]
L.nann_debug_tc_trace(C.c_void_p(buf.data_ptr()))
nb.score_ids(sc, u, emb_d, ids)
L.nann_debug_tc_trace(None)
t = buf.cpu().numpy().reshape(64, 48)
names = {0: "mma tile start", 1: "mma x units issued (d1_full committed)", 2: "mma d2_empty ok", 3: "mma unit 2 issued", 4: "mma unit 3 issued",
         5: "mma unit 4 issued", 6: "mma unit 5 issued", 7: "mma d2 committed", 8: "epi iter start", 9: "epi d1_full ok", 10: "epi epi1 done",
         11: "epi gather(next) loads issued", 12: "epi x(next) written", 13: "epi d2_full ok", 14: "epi tile done"}
for tile in (10, 11):
    base = t[tile][0]
    print(f"--- tile {tile} (cycles from event 0)")
    for ev in sorted(names, key=lambda e: t[tile][e]):
        if t[tile][ev]: print(f"{t[tile][ev]-base:9d}  {names[ev]}")
b11 = t[11][0]
if True:
    print("tile 11 units: (a_full ok, hi stage ok, lo stage ok, all issued) from tile start")
    for u in range(10):
        print(f"   unit {u}:", [int(x - b11) for x in t.reshape(-1)[60 * 48 + 4 * u:60 * 48 + 4 * u + 4]])
print("tile 11 ring stages: producer saw slot empty:", [int(x - b11) for x in t[62][:40]])
print("tile 11 ring stages: mma saw stage full:   ", [int(x - b11) for x in t[63][:40]])
print("mma warp waits per tile: ring", [int(t[i][46]) for i in range(8, 12)], "a_full", [int(t[i][47]) for i in range(8, 12)])
print("tile period (cycles):", [int(t[i+1][0]-t[i][0]) for i in range(5, 15)])
