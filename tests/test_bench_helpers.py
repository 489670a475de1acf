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


def test_blocked_corpus_is_sliceable_at_any_boundary(monkeypatch):
    """corpora above bench.BIG are DEFINED block-wise: any [lo, hi) slice must equal the same rows of a larger slice"""
    sys.path.insert(0, ROOT)
    import bench
    monkeypatch.setattr(bench, "BIG", 1000)
    monkeypatch.setattr(bench, "N_BLOCKS", 8)
    n = 4000                                             # 8 blocks of 500 rows
    full_e, full_i = bench.corpus_rows(n, 0, n)
    assert full_e.shape == (n, 128) and len(set(full_i.tolist())) == n
    for lo, hi in ((0, 500), (250, 1250), (1999, 2001), (3500, 4000)):
        e, i = bench.corpus_rows(n, lo, hi)
        np.testing.assert_array_equal(e, full_e[lo:hi])
        np.testing.assert_array_equal(i, full_i[lo:hi])
    assert bench.query_pool(n).shape == (500, 128)


def test_per_beam_shard_scales():
    from nann_b200.distributed import shard_beams, shard_level_topn
    T = [100, 200, 200, 200, 200, 200]
    assert shard_beams(T, 1, [3, 3, 3, 3, 3]) == T
    assert shard_beams(T, 8, [1.0] * 5) == shard_level_topn(T, 8)
    t = shard_beams(T, 8, [1.0837, 1.275, 1.5, 1.5, 0.2953])
    assert t == [14, 32, 38, 38, 8, 50]                  # the beams of the N = 8 and 100M runs (profiles/r02_bench_8gpu_sharded.json)
    assert all(x >= 8 for x in t[:5]) and t[5] <= sum(t[1:5]) and 8 * t[5] >= T[5]


def _lines(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return [json.loads(l) for l in f if l.strip().startswith("{")]


def test_committed_bench_lines_carry_the_whole_contract():
    """the bench lines kept under profiles/ (what DESIGN.md quotes) have every key of the bench contract, and the numbers
    inside are consistent with each other"""
    base = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline")
    one = _lines("r02_bench_1gpu.json")[-1]
    for k in base:
        assert k in one, k
    assert one["n_gpus"] == 1 and one["higher_is_better"] is True and one["vs_baseline"] is None and one["data"] == "synthetic"
    assert "workload" in one["config"] and "configs[1]" in one["config"]["workload"] and "l2" in one["config"]
    assert one["gpu_launches"] > 0 and one["warmup"] >= 3
    # value = batch * steps / device time
    assert abs(one["value"] - 256 * one["steps"] / (one["ms_per_step"] * one["steps"] / 1e3)) < 1e-6 * one["value"]
    e = one["e2e"]
    assert e["h2d_bytes_per_step"] == 256 * 128 * 4 and e["d2h_bytes_per_step"] > 256 * 200 * 12 and 0 < e["value"] <= one["value"] * 1.02
    r = one["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
    assert abs(r["achieved"] - r["rows_per_launch"] * r["algorithmic_flops_per_row"] / (r["avg_launch_ms"] / 1e3) / 1e12) < 1e-6 * r["achieved"]
    c = one["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["unit"] == one["unit"] and c["sample"] and c["value"] > 0
    assert c["exact_path_ids_equal_to_cpu"] and c["exact_path_scores_bit_equal"]        # parity side check of the bench itself
    assert set(one["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(one["clocks"]["reasons"]))
    ref = _lines("r02_bench_reference.json")[-1]
    assert ref["impl"] == "reference" and ref["config"]["workload"] == one["config"]["workload"] and ref["metric"] == one["metric"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["e2e"]["value"] == ref["value"] and ref["cpu_baseline"]["kind"] == "port"
    for n in (2, 4, 8):
        d = _lines(f"r02_bench_{n}gpu_dist.json")[-1]
        for k in base:
            assert k in d, (n, k)
        assert d["n_gpus"] == n and d["scaling"] == "weak" and d["ids_bit_identical_to_one_gpu_search"] is True and d["recall_held"] is True
        assert d["gpu_launches"] > 0
        s = _lines(f"r02_bench_{n}gpu_sharded.json")[-1]
        assert s["replica_mode"]["value"] > 0            # the replica curve of the same box
        assert s["n_gpus"] == n and s["recall_held"] is True and abs(s["recall_at_k_vs_bruteforce"] - s["recall_target"]) <= 0.005 + 1e-9
