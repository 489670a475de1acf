"""Host-side pieces of bench.py that run without a GPU: the sharded corpus definition, per-shard beam widths and the
stdout contract (exactly one JSON line)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_corpus_blocks_are_disjoint_and_deterministic():
    sys.path.insert(0, ROOT)
    import bench
    e0, i0 = bench.corpus_block(1000, 0)
    e1, i1 = bench.corpus_block(1000, 1)
    e0b, i0b = bench.corpus_block(1000, 0)
    np.testing.assert_array_equal(e0, e0b)
    np.testing.assert_array_equal(i0, i0b)
    assert e0.shape == (1000, 128) and not np.array_equal(e0, e1)
    assert sorted(i0.tolist()) == list(range(0, 1000)) and sorted(i1.tolist()) == list(range(1000, 2000))   # global ids, permuted
    np.testing.assert_allclose(np.linalg.norm(e1, axis=1), 1.0, rtol=1e-5)


def test_shard_level_topn_keeps_every_topk_feasible():
    from nann_b200.distributed import shard_level_topn
    for T in ([100, 200, 200, 200, 200, 200], [100, 200, 400, 400, 400, 200]):
        assert shard_level_topn(T, 1) == T
        for G in (2, 4, 8):
            t = shard_level_topn(T, G)
            assert all(a >= 8 for a in t[:5]) and t[5] <= sum(t[1:5])       # the final top-k finds its k results
            assert G * t[5] >= T[5]                                          # the merge has at least k candidates
            assert all(a <= b for a, b in zip(t[:5], T[:5]))


def test_reference_arm_prints_exactly_one_json_line(tmp_path):
    """`bench.py --impl reference` on a corpus small enough for the CPU builder; stdout must be one JSON object with
    the contract's keys (libraries may print banners: those have to land on stderr)."""
    env = dict(os.environ, NANN_BENCH_CACHE=str(tmp_path), CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--n-items", "110000", "--cpu-sample", "4"], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
