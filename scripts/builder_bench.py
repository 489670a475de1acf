"""times nann_hnsw_build on a synthetic corpus resident in HBM:  python scripts/builder_bench.py N [reps]"""
import sys
import time

import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from nann_b200 import builder  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
g = torch.Generator(device="cuda").manual_seed(0)
e = torch.randn((n, 128), generator=g, device="cuda")
e /= e.norm(dim=1, keepdim=True)
torch.cuda.synchronize()
for rep in range(reps):
    t = time.time()
    out = builder.build_hnsw(e, M=32, seed=4, return_stats=True)
    print(f"builder n={n}: {time.time() - t:.2f}s", out["stats"], [len(v) for v in out["values"]], flush=True)
