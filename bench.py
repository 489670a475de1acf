#!/usr/bin/env python
"""bench.py -- queries/s of the model-scored HNSW retrieval hot path (BASELINE.json metric).

A "step" = one batch of queries through the whole exec.pb dataflow (5 scoring rounds, 4 expand +
visited-filter rounds, 6 top-k) on synthetic data.  Default workload = BASELINE configs[1]:
1M items d=128 f32, batch=256 queries, scoring MLP 2x512, ef_search=200
(level_topn=[100,200,200,200,200,200]), one B200.

  python bench.py --gpus 1 --steps 10 --warmup 3            # this repo's CUDA path
  python bench.py --impl reference ...                       # the reference algorithm on the host cores
  torchrun --nproc-per-node N bench.py --gpus N ...          # the SAME --n-items corpus row-sharded N ways,
                                                             # global batch N x --batch, per-shard beams calibrated
                                                             # so that recall@200 equals the one-GPU operating point
  ... --n-items 100000000 --gpus 8                           # BASELINE configs[3]

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import hashlib
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# one hardware work queue per stream (default 8 are shared round-robin): the shard group's merge stream holds a kernel
# that waits for the peers' deliveries, and work of another stream queued behind it in the SAME hardware queue would
# wait too (no deadlock across GPUs, but the search/merge overlap would be lost)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

EF_TOPN = {200: [100, 200, 200, 200, 200, 200],      # reference README.md:216
           400: [100, 200, 400, 400, 400, 200]}      # NANN_impls/nann/benchmark/gen_runmeta.py:23
MAC_PER_ROW = 256 * 512 + 512 * 512 + 512            # un-hoisted algorithmic count
ROW_BYTES = 512
CACHE = os.environ.get("NANN_BENCH_CACHE", "/tmp/nann_b200_bench_cache")
BIG = 16_000_000                                     # corpora above this are generated block-wise
N_BLOCKS = 64
SCALES = (0.5, 0.625, 0.75, 0.875, 1.0, 1.125, 1.25, 1.5, 1.75, 2.0, 2.5, 3.0, 4.0)
RECALL_TOL = 0.005


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def _by_path(name):
    """nann_b200/<name>.py as a plain module, WITHOUT importing the package: the reference arm must not load
    libnann_b200.so (these modules are numpy / torch only)."""
    key = "_nann_helper_" + name
    if key in sys.modules:
        return sys.modules[key]
    spec = importlib.util.spec_from_file_location(key, os.path.join(ROOT, "nann_b200", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules[key] = m
    spec.loader.exec_module(m)
    return m


def nix():
    return _by_path("index")


def shard_bounds(n, world, rank):
    return _by_path("distributed").shard_bounds(n, world, rank)


def shard_topn(T, world, scale=1.0):
    return _by_path("distributed").shard_level_topn(T, world, scale)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d.get("hbm_gbs", 6650.0), tf_burst=d.get("bf16_tflops", 1590.0),
                    tf_sustained=d.get("bf16_tflops_sustained", 1400.0), source="measured")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------------------
# workload: ONE corpus of --n-items rows whatever the number of GPUs; rank r of N searches rows [lo, hi)
# ------------------------------------------------------------------------------------------------
def corpus_block(n_rows, block):
    """rows and GLOBAL item ids of block `block` of a block-wise defined corpus (seed 100+block)"""
    emb = nix().synthetic_corpus(n_rows, 128, seed=100 + block)
    ids = np.int64(block) * n_rows + nix().synthetic_item_ids(n_rows, seed=200 + block)
    return emb, ids


def corpus_rows(n_items, lo, hi):
    """rows [lo, hi) of the corpus and their item ids.  Up to 16M rows the corpus is one seeded draw (the same
    arrays for every N, identical to the single-GPU configs); above, it is DEFINED as 64 seeded blocks so that
    no rank ever materialises more than the blocks its rows touch."""
    if n_items <= BIG:
        full = nix().synthetic_corpus(n_items, 128, seed=0)
        ids = nix().synthetic_item_ids(n_items, seed=1)
        return np.ascontiguousarray(full[lo:hi]), np.ascontiguousarray(ids[lo:hi])
    per = -(-n_items // N_BLOCKS)
    embs, idl = [], []
    for b in range(lo // per, (hi - 1) // per + 1):
        b_lo, b_hi = b * per, min((b + 1) * per, n_items)
        e, _ = corpus_block(b_hi - b_lo, b)
        s0, s1 = max(lo, b_lo) - b_lo, min(hi, b_hi) - b_lo
        embs.append(e[s0:s1])
        idl.append(np.arange(b_lo + s0, b_lo + s1, dtype=np.int64) + 7_000_000_000)     # id != row
    return np.ascontiguousarray(np.concatenate(embs)), np.concatenate(idl)


def query_pool(n_items):
    """the rows queries are drawn from (+ noise): the whole corpus, or block 0 of a block-wise corpus"""
    if n_items <= BIG:
        return nix().synthetic_corpus(n_items, 128, seed=0)
    per = -(-n_items // N_BLOCKS)
    return corpus_block(min(per, n_items), 0)[0]


def builder_name(device):
    """the hand-written CUDA builder of the library when a GPU is there (nann_b200/builder.py), the torch builder
    otherwise (CPU: small test corpora only)"""
    if str(device).startswith("cuda") and os.environ.get("NANN_BENCH_TORCH_BUILDER", "0") != "1" and \
            os.path.exists(os.path.join(ROOT, "nann_b200", "builder.py")):
        return "cuda-builder"
    return "torch-builder"


def build_graph(emb, seed, device):
    if builder_name(device) == "cuda-builder":
        from nann_b200 import builder
        dev_index = int(str(device).split(":")[1]) if ":" in str(device) else 0
        return builder.build_hnsw(emb, M=32, start_level=2, seed=seed, device=dev_index,
                                  values_dtype=np.int32 if emb.shape[0] > 4_000_000 else np.int64)
    return nix().build_hnsw(emb, M=32, start_level=2, seed=seed, device=device)


def shard_dir(n_items, world, rank):
    key = hashlib.sha1(json.dumps([n_items, world, rank, 128, 32, 4, "r2"]).encode()).hexdigest()[:16]
    return os.path.join(CACHE, f"shard_{n_items}_{world}_{rank}_{key}")


def get_shard(n_items, world, rank, device):
    """-> dict(emb, item_ids (GLOBAL ids), ep, values, row_splits) for rows [lo, hi) of the corpus."""
    lo, hi = shard_bounds(n_items, world, rank)
    root = shard_dir(n_items, world, rank)
    embs_dir, index_dir = os.path.join(root, "embeddings"), os.path.join(root, "index")
    big = hi - lo > 4_000_000        # ~10 GB of files per shard: keep them in memory
    if big or not os.path.exists(os.path.join(root, "done")):
        t = time.time()
        emb, ids = corpus_rows(n_items, lo, hi)
        t1 = time.time()
        g = build_graph(emb, 4 + rank, device)
        log(f"[bench] rank {rank}: rows [{lo},{hi}) generated in {t1 - t:.1f}s, HNSW built in {time.time() - t1:.1f}s ({builder_name(device)}, {device})")
        if big:
            return dict(emb=emb, item_ids=ids, ep=g["enter_points"], values=g["values"], row_splits=g["row_splits"])
        nix().save_index(embs_dir, index_dir, emb, ids, g)
        open(os.path.join(root, "done"), "w").write("ok")
    emb, item_ids, g = nix().load_index_arrays(embs_dir, index_dir)
    return dict(emb=emb, item_ids=item_ids, ep=g["enter_points"], values=g["values"], row_splits=g["row_splits"])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 3.0:      # nvidia-smi needs ~0.1-1 s to print its first row
                time.sleep(0.01)
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.perf_counter()] + [c.strip() for c in line.split(",")])

    def window(self, t0, t1):
        """keep the samples taken inside [t0, t1] (perf_counter); if the window is shorter than the sampling
        period, keep the samples of the whole sampler lifetime (warm-up steps = the same load)"""
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        self.scope = "timed region" if inside else "warm-up + timed region"
        if inside:
            self.rows = inside

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.12)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            r = r[1:]
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "scope": getattr(self, "scope", "sampler lifetime")}


def nvlink_tx_rx_kib(gpu_index):
    """NVLink data counters of GPU `gpu_index`, summed over its links, in KiB: (tx, rx, source) or None.
    NVML field values first (NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX/RX, scope = all links), nvidia-smi as a fallback;
    on driver stacks that do not expose the counters both say N/A."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        vals = pynvml.nvmlDeviceGetFieldValues(h, [(pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, 0xFFFFFFFF),
                                                   (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, 0xFFFFFFFF)])
        if all(v.nvmlReturn == 0 for v in vals):
            return int(vals[0].value.ullVal), int(vals[1].value.ullVal), "nvml field values"
    except Exception:
        pass
    try:
        r = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(gpu_index)], capture_output=True, text=True, timeout=10)
        tx = rx = 0
        seen = False
        for line in r.stdout.splitlines():
            parts = line.replace(":", " ").split()
            if "Tx" in parts and "KiB" in parts:
                tx += int(parts[parts.index("KiB") - 1]); seen = True
            if "Rx" in parts and "KiB" in parts:
                rx += int(parts[parts.index("KiB") - 1]); seen = True
        return (tx, rx, "nvidia-smi nvlink -gt d") if seen else None
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
def workload_config(args, T, world):
    which = {(1_000_000, 256, 200): "BASELINE configs[1]", (10_000_000, 1024, 400): "BASELINE configs[2]"}.get(
        (args.n_items, args.batch, args.ef), "BASELINE configs[4] sweep point" if args.n_items == 1_000_000 else "custom")
    l2 = "inputs larger than L2: 512 MB embedding table + 260 MB graph per 1M rows, fresh queries every step"
    if world > 1:
        shape = "BASELINE configs[3]" if args.n_items == 100_000_000 and world == 8 else f"BASELINE configs[3] shape on the {which} corpus"
        return {"workload": f"{args.n_items} items d=128 f32 row-sharded over {world} GPUs ({-(-args.n_items // world)} rows + their own HNSW each), "
                            f"global batch={args.batch * world} queries ({args.batch} per GPU), scoring MLP 2x512, ef_search={args.ef}, HNSW M=32 ({shape})",
                "level_topn": list(T),
                "parallelism": f"corpus row-sharded x{world}: every query visits every shard; per-shard top-k pushed into every rank's window over "
                               f"NVLink by the final top-k kernel, merge kernel per rank (nann_search_sharded); batch grows with N (weak scaling)",
                "l2": l2}
    return {"workload": f"{args.n_items} items d=128 f32, batch={args.batch} queries, scoring MLP 2x512, "
                        f"ef_search={args.ef}, HNSW M=32 ({which})",
            "level_topn": list(T), "parallelism": "1 GPU", "l2": l2}


def run_reference(args, T, rank, world):
    """The reference algorithm (CPU custom-op path restated in oracle/) on the host cores: batch=1 per request, one
    in-flight request per core (blaze-benchmark consumers).  This process loads oracle/ only -- the HNSW files it
    searches are produced by the offline builder in a separate process when they are not cached yet."""
    if rank != 0:
        return None
    from oracle import oracle as orc
    sw = _by_path("scorer_weights")
    key_done = os.path.join(shard_dir(args.n_items, 1, 0), "done")
    if args.n_items > 200_000 and not os.path.exists(key_done):
        # index construction is offline tooling, not the timed path: run it out of process
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--build-index-only", "--n-items", str(args.n_items)],
                           stdout=subprocess.DEVNULL, stderr=sys.stderr)
        if r.returncode != 0:
            log("[bench] out-of-process index build failed; building on the CPU")
    sh = get_shard(args.n_items, 1, 0, "cpu")
    cores = os.cpu_count() or 1
    oix = orc.Index(sh["emb"], sh["item_ids"], sh["ep"].astype(np.int32), [v.astype(np.int32) for v in sh["values"]], sh["row_splits"])
    om = orc.Mlp(*sw.mlp_weights(seed=3))
    sample_q = args.cpu_sample or int(min(args.batch, max(32, 4 * cores)))
    queries = nix().synthetic_queries(sh["emb"], sample_q * (args.steps + args.warmup), seed=2)
    for w in range(args.warmup):
        oix.search_batch_mlp(om, queries[w * sample_q:(w + 1) * sample_q], T, nthreads=cores)
    secs, rows = 0.0, 0
    for s in range(args.steps):
        o = (args.warmup + s) * sample_q
        r = oix.search_batch_mlp(om, queries[o:o + sample_q], T, nthreads=cores)
        secs += r["seconds"]; rows += r["n_scored"]
        assert np.all(r["status"] == 0)
    qps = sample_q * args.steps / secs
    sample = f"{sample_q} queries/step (of the {args.batch}-query batch), {args.steps} steps, one request per core"
    return {"impl": "reference", "metric": "queries/sec at fixed recall@200", "value": qps, "unit": "queries/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * secs / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, T, 1),
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample,
                             "rows_scored_per_query": rows / (sample_q * args.steps)},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


# ------------------------------------------------------------------------------------------------
class Truth:
    """brute-force top-k under the same scorer over the WHOLE corpus for the evaluation queries: every rank scores
    its own rows with the product's scorer (nann_blaze_xla_run, device in/out), the per-shard top-k are gathered with
    torch.distributed and merged on every rank (evaluation plumbing, outside the timed region)."""

    def __init__(self, nb, sc, emb_dev, item_ids_dev, queries, k, world, dev):
        import torch
        import torch.distributed as dist
        n = emb_dev.shape[0]
        out = torch.empty(n, dtype=torch.float32, device=dev)
        kk = min(k, n)
        loc_s = torch.empty((len(queries), kk), dtype=torch.float32, device=dev)
        loc_i = torch.empty((len(queries), kk), dtype=torch.int64, device=dev)
        for q in range(len(queries)):
            nb.blaze_xla_op(sc, queries[q], emb_dev, out=out)
            s, i = torch.topk(out, kk)
            loc_s[q], loc_i[q] = s, item_ids_dev[i]
        if world > 1:
            gs = [torch.empty_like(loc_s) for _ in range(world)]
            gi = [torch.empty_like(loc_i) for _ in range(world)]
            dist.all_gather(gs, loc_s); dist.all_gather(gi, loc_i)
            cs, ci = torch.cat(gs, 1), torch.cat(gi, 1)
            s, o = torch.topk(cs, k, dim=1)
            loc_i = torch.gather(ci, 1, o)
        self.ids = loc_i.cpu().numpy()
        self.k = k

    def recall(self, got_ids):
        n = min(len(got_ids), len(self.ids))
        hits = sum(len(set(self.ids[q].tolist()) & set(np.asarray(got_ids[q]).tolist())) for q in range(n))
        return hits / (n * self.k)


def run_b200(args, T, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import nann_b200 as nb
    from nann_b200 import scorer_weights as sw
    from nann_b200.distributed import ShardGroup
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    k = T[5]
    sh = get_shard(args.n_items, world, rank, dev)
    ix = nb.Index.from_arrays(sh["emb"], sh["item_ids"], sh["ep"], sh["values"], sh["row_splits"], device=local_rank)
    sc = nb.Scorer.mlp(*sw.mlp_weights(seed=3), device=local_rank)
    if args.precision == "tensor":
        sc.set_precision(nb.SCORER_TENSOR)
    B = args.batch * world          # N > 1: the global batch grows with N and every shard sees all of it
    n_steps = args.warmup + args.steps
    pool = query_pool(args.n_items)                     # every rank draws the same queries
    queries = nix().synthetic_queries(pool, B * n_steps, seed=2)
    n_eval = max(0, args.eval_queries)
    eval_q = nix().synthetic_queries(pool, max(n_eval, 1), seed=5)
    del pool
    stream = torch.cuda.Stream(device=dev)              # a non-blocking stream (the legacy default stream serialises with every other stream)
    extra = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- recall bookkeeping (outside the timed region): brute-force truth, the one-GPU operating point, calibration
    truth = None
    if n_eval > 0:
        t0 = time.time()
        emb_dev = torch.from_numpy(sh["emb"]).to(dev)           # evaluation copy (the index owns its own)
        truth = Truth(nb, sc, emb_dev, torch.from_numpy(sh["item_ids"]).to(dev), eval_q, k, world, dev)
        del emb_dev
        torch.cuda.empty_cache()
        log(f"[bench] rank {rank}: brute-force truth for {n_eval} queries in {time.time() - t0:.1f}s")

    grp, exchange = None, "none (1 GPU)"
    Ts, scale = list(T), 1.0
    if world > 1:
        k_s_max = min(k, 4096 // world * 2)
        grp = ShardGroup(rank, world, max(B, n_eval), k_s_max, device=local_rank)
        try:
            grp.connect_torch()
            exchange = "in-library: final top-k kernel stores (score,id) records into every rank's window over NVLink (CUDA IPC peer mappings), merge kernel on the group's stream"
        except Exception as e:                                   # no peer mappings in this container: NCCL transport
            log(f"[bench] rank {rank}: peer windows unavailable ({e!r}); falling back to torch.distributed all_gather")
            grp = None
            exchange = f"torch.distributed all_gather + nann_merge_topk (peer windows unavailable: {str(e)[:120]})"
        ok = torch.tensor([1.0 if grp is not None else 0.0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0.0 and grp is not None:                # all ranks must use the same transport
            grp = None
            exchange = "torch.distributed all_gather + nann_merge_topk (a peer could not map the windows)"

    def sharded_host(se, users_np, Ts_):
        """host queries in, merged host ids out, every rank the same result"""
        if grp is not None:
            sc_, ids_, st_ = grp.search(se, users_np, Ts_, k)
            return ids_, sc_, st_
        from nann_b200 import distributed as nd
        u = torch.from_numpy(users_np).to(dev)
        o_i = torch.empty((len(users_np), Ts_[5]), dtype=torch.int64, device=dev)
        o_s = torch.empty((len(users_np), Ts_[5]), dtype=torch.float32, device=dev)
        m_sc, m_id, st_, _ = nd.sharded_search(se, u, Ts_, k, o_i, o_s, nb.merge_topk)
        return m_id, m_sc, st_

    recall_target = None
    if world > 1 and truth is not None and args.shard_scale <= 0 and not args.shard_scales:
        from nann_b200 import distributed as nd
        # a second, independent query set drives the calibration; the recall that is REPORTED (and checked against the
        # one-GPU operating point) is measured on eval_q afterwards
        cal_q = nix().synthetic_queries(query_pool(args.n_items), n_eval, seed=6)
        emb_dev = torch.from_numpy(sh["emb"]).to(dev)
        truth_cal = Truth(nb, sc, emb_dev, torch.from_numpy(sh["item_ids"]).to(dev), cal_q, k, world, dev)
        del emb_dev
        torch.cuda.empty_cache()
        # the one-GPU operating point: rank 0 searches the UNSHARDED index with the full beams
        cal_target = None
        if args.n_items <= BIG:
            rt = torch.zeros(3, dtype=torch.float64, device=dev)
            if rank == 0:
                t0 = time.time()
                sh1 = get_shard(args.n_items, 1, 0, dev)
                ix1 = nb.Index.from_arrays(sh1["emb"], sh1["item_ids"], sh1["ep"], sh1["values"], sh1["row_splits"], device=local_rank)
                se1 = nb.Searcher(ix1, sc, n_eval, T)
                r1 = se1.search(eval_q[:n_eval], T)
                rt[0] = truth.recall(r1["ids"])
                rt[1] = truth_cal.recall(se1.search(cal_q, T)["ids"])
                rt[2] = float(r1["n_scored"].sum()) / n_eval
                del se1, ix1, sh1
                log(f"[bench] one-GPU operating point: recall@{k} = {rt[0].item():.4f} ({rt[1].item():.4f} on the calibration queries), "
                    f"{rt[2].item():.0f} rows scored per query ({time.time() - t0:.1f}s)")
            dist.broadcast(rt, 0)
            recall_target, cal_target = rt[0].item(), rt[1].item()
            extra["one_gpu_rows_scored_per_query"] = rt[2].item()
        cal_dev = torch.from_numpy(cal_q).to(dev)
        trials = []

        def evaluate(scales):
            """-> (recall on the calibration queries, rows scored per query summed over the shards)"""
            Ts_ = nd.shard_beams(T, world, scales)
            se_ = nb.Searcher(ix, sc, n_eval, Ts_)
            o_i = torch.empty((n_eval, Ts_[5]), dtype=torch.int64, device=dev)
            o_s = torch.empty((n_eval, Ts_[5]), dtype=torch.float32, device=dev)
            _, m_id, _, stats = nd.sharded_search(se_, cal_dev, Ts_, k, o_i, o_s, nb.merge_topk)
            rows_ = torch.tensor([float(stats["n_scored"].sum())], dtype=torch.float64, device=dev)
            dist.all_reduce(rows_)
            rec = truth_cal.recall(np.asarray(m_id))
            trials.append({"scales": [round(float(x), 4) for x in scales], "shard_level_topn": Ts_, "recall": rec,
                           "rows_scored_per_query": rows_.item() / n_eval})
            return rec, rows_.item() / n_eval

        # (1) smallest uniform scale that holds the target (no unsharded operating point to hold: plain T / world)
        scales = [1.0] * 5
        if cal_target is None:
            evaluate(scales)
        else:
            for s_ in SCALES:
                rec, _ = evaluate([s_] * 5)
                scales = [s_] * 5
                if rec >= cal_target - RECALL_TOL:
                    break
        # (2) lower the beams one by one while the target still holds (each accepted step scores fewer rows)
        if cal_target is not None:
            improved = True
            while improved and len(trials) < 60:
                improved = False
                for l in range(5):
                    cand = list(scales)
                    cand[l] = cand[l] * 0.85
                    if nd.shard_beams(T, world, cand) == nd.shard_beams(T, world, scales):
                        continue
                    rec, _ = evaluate(cand)
                    if rec >= cal_target - RECALL_TOL:
                        scales, improved = cand, True
        # (3) check on the report queries; widen everything a notch if the independent set disagrees
        for _ in range(6):
            Ts = nd.shard_beams(T, world, scales)
            se_ = nb.Searcher(ix, sc, n_eval, Ts)
            rec = truth.recall(sharded_host(se_, eval_q[:n_eval], Ts)[0])
            del se_
            if recall_target is None or rec >= recall_target - RECALL_TOL:
                break
            scales = [x * 1.06 for x in scales]
        scale = [round(float(x), 4) for x in scales]
        extra["calibration"] = trials
        del cal_dev
    elif world > 1:
        from nann_b200 import distributed as nd
        scale = [float(x) for x in args.shard_scales.split(",")] if args.shard_scales else [args.shard_scale if args.shard_scale > 0 else 1.0] * 5
        Ts = nd.shard_beams(T, world, scale)

    se = nb.Searcher(ix, sc, B, Ts)
    k_s = Ts[5]
    # batches in flight (one GPU): searcher j + stream j serve steps j, j+P, ...; every batch is still one whole call
    P = max(1, args.inflight) if world == 1 else 1
    ses = [se] + [nb.Searcher(ix, sc, B, Ts) for _ in range(P - 1)]
    streams = [stream] + [torch.cuda.Stream(device=dev) for _ in range(P - 1)]
    q_dev = torch.from_numpy(queries).to(dev)
    q_pin = torch.from_numpy(queries).pin_memory()
    outs = [(torch.empty((B, k), dtype=torch.int64, device=dev), torch.empty((B, k), dtype=torch.float32, device=dev),
             torch.empty((B,), dtype=torch.int32, device=dev)) for _ in range(max(2, P))]
    if world > 1 and grp is None:
        from nann_b200 import distributed as nd
        ids_d = torch.empty((B, k_s), dtype=torch.int64, device=dev)
        sc_d = torch.empty((B, k_s), dtype=torch.float32, device=dev)

    def step_device(i):
        """inputs resident in HBM, results left in HBM, nothing synchronises the host"""
        u = q_dev[i * B:(i + 1) * B]
        if world == 1:
            ses[i % P].search_async(u, Ts, *outs[i % len(outs)], stream=streams[i % P])
        elif grp is not None:
            grp.search(se, u, Ts, k, *outs[i & 1], stream=stream)
        else:
            nd.sharded_search(se, u, Ts, k, ids_d, sc_d, nb.merge_topk, stream=stream)

    p_ids = torch.empty((B, k_s), dtype=torch.int64, device=dev)
    p_sc = torch.empty((B, k_s), dtype=torch.float32, device=dev)

    def step_profile(i):
        """the same step (this rank's shard search) through the call that returns the per-stage CUDA-event clocks;
        it synchronises every step, which is why it is not the pass `value` is taken from"""
        se.search_device(q_dev[i * B:(i + 1) * B], Ts, p_ids, p_sc, stream=stream)

    def step_e2e(i):
        # host buffers in, host results out: H2D of the queries and D2H of ids+scores inside the timed region
        if world == 1:
            return ses[i % P].search(q_pin[i * B:(i + 1) * B].numpy(), Ts, stream=streams[i % P])
        return sharded_host(se, q_pin[i * B:(i + 1) * B].numpy(), Ts)

    step_wall = []

    def timed(fn, profile, drain=False, threads=1):
        for w in range(args.warmup):
            fn(w)
        for j in range(P if P > 1 else 0):                  # every searcher in flight has run at least twice
            fn(j)
        if drain and grp is not None:
            grp.wait()
        barrier()
        se.set_profile(profile)
        l0 = nb.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        del step_wall[:]
        t0 = time.perf_counter()
        timed.window = [t0, t0]
        e0.record(stream)

        def one(s):
            ts = time.perf_counter()
            fn(args.warmup + s)
            step_wall.append(time.perf_counter() - ts)

        if threads > 1:        # synchronous public calls from `threads` host threads, thread j = searcher j = steps j, j+P, ...
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(threads) as pool:
                for f in [pool.submit(lambda j=j: [one(s) for s in range(j, args.steps, threads)]) for j in range(threads)]:
                    f.result()
        else:
            for s in range(args.steps):
                one(s)
        if drain and grp is not None:
            grp.wait(stream=stream, host_block=False)      # the timed region ends when the last merge has run
        for st_ in streams[1:]:                             # ... and when every stream in flight has drained
            ev_ = torch.cuda.Event()
            ev_.record(st_)
            stream.wait_event(ev_)
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        timed.window[1] = t0 + wall
        dev_ms = e0.elapsed_time(e1)
        prof = se.profile() if profile else None
        se.set_profile(False)
        t = torch.tensor([dev_ms, wall * 1000.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item(), t[1].item(), nb.launch_count() - l0, prof

    nvl0 = nvlink_tx_rx_kib(local_rank) if (world > 1 and rank == 0) else None
    with ClockSampler(local_rank) as clk:
        dev_ms, _, launches, _ = timed(step_device, False, drain=True)
        clk.window(*timed.window)
    nvl1 = nvlink_tx_rx_kib(local_rank) if (world > 1 and rank == 0) else None
    # stage clocks (scorer launches for the roofline): the same steps with CUDA events around every stage
    prof_ms, _, _, prof = timed(step_profile, True)
    e2e_dev_ms, e2e_wall_ms, _, _ = timed(step_e2e, False, threads=P)
    # the public call is synchronous (host results are returned), so a step's wall time is the batch latency
    lat = sorted(step_wall)
    latency_ms = {"p50": 1000.0 * lat[len(lat) // 2], "max": 1000.0 * lat[-1], "batch": B,
                  "what": "wall time of one public API call (host queries in, host ids+scores out)" +
                          (f", {P} such calls in flight on the GPU" if P > 1 else "")}

    # ---- recall@k of the measured configuration (outside the timed region)
    try:
        if truth is None:
            raise RuntimeError("recall evaluation disabled (--eval-queries 0)")
        se_e = nb.Searcher(ix, sc, n_eval, Ts) if n_eval > B else se
        got = se_e.search(eval_q[:n_eval], Ts)["ids"] if world == 1 else sharded_host(se_e, eval_q[:n_eval], Ts)[0]
        extra["recall_at_k_vs_bruteforce"] = truth.recall(got)
        extra["recall_queries"] = n_eval
        if world > 1:
            extra["recall_target"] = recall_target
            extra["recall_target_source"] = ("the unsharded index searched with the full beams on rank 0, same queries, this run"
                                             if recall_target is not None else "none (corpus too large for one unsharded index build here)")
            extra["recall_held"] = bool(recall_target is not None and extra["recall_at_k_vs_bruteforce"] >= recall_target - RECALL_TOL)
            extra["shard_beam_scale"] = scale
    except Exception as e:  # never lose the bench line over the side measurements
        extra["recall_error"] = repr(e)[:200]

    # ---- replica mode (SURVEY 8e alternative; what the reference's session pool does): the WHOLE index on every GPU,
    # --batch queries per GPU, no exchange
    if world > 1 and args.n_items <= BIG and not args.no_replica:
        try:
            if rank == 0:
                get_shard(args.n_items, 1, 0, dev)               # cached by the recall-target pass; builds otherwise
            barrier()
            sh1 = get_shard(args.n_items, 1, 0, dev)
            ix1 = nb.Index.from_arrays(sh1["emb"], sh1["item_ids"], sh1["ep"], sh1["values"], sh1["row_splits"], device=local_rank)
            se1 = nb.Searcher(ix1, sc, args.batch, T)
            o1 = [(torch.empty((args.batch, k), dtype=torch.int64, device=dev), torch.empty((args.batch, k), dtype=torch.float32, device=dev))
                  for _ in range(2)]
            b1 = args.batch

            def step_replica(i):       # rank r serves its own slice of the global batch
                u = q_dev[i * B + rank * b1:i * B + (rank + 1) * b1]
                se1.search_async(u, T, *o1[i & 1], stream=stream)

            r_ms, _, _, _ = timed(step_replica, False)
            extra["replica_mode"] = {"value": B * args.steps / (r_ms / 1000.0), "unit": "queries/s", "ms_per_step": r_ms / args.steps,
                                     "what": f"whole {args.n_items}-row index on every GPU, {b1} queries per GPU and step, no exchange "
                                             f"(the reference's replica model, blaze-benchmark/benchmark/core/model.cc:192-234); recall = the one-GPU operating point"}
            del se1, ix1, sh1
        except Exception as e:
            extra["replica_mode"] = {"error": repr(e)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:   # cpu_baseline: rank 0 at N=1 only
        try:
            from oracle import oracle as orc
            cores = os.cpu_count() or 1
            oix = orc.Index(sh["emb"], sh["item_ids"], sh["ep"].astype(np.int32), [v.astype(np.int32) for v in sh["values"]], sh["row_splits"])
            om = orc.Mlp(*sw.mlp_weights(seed=3))
            n_warm = int(min(len(queries), max(32, 2 * cores)))
            rw = oix.search_batch_mlp(om, queries[:n_warm], T, nthreads=cores)   # warm + rate estimate
            # bounded sample: about 12 s of host work on this box's cores (same queries the GPU arm is timed on)
            sample_q = args.cpu_sample or int(min(len(queries), max(64, 12.0 * n_warm / max(rw["seconds"], 1e-3))))
            r = oix.search_batch_mlp(om, queries[:sample_q], T, nthreads=cores)
            cpu = {"value": sample_q / r["seconds"], "unit": "queries/s", "cores": cores, "kind": "port",
                   "sample": f"first {sample_q} queries of the run, one request per core (batch=1 each), {r['seconds']:.1f}s",
                   "rows_scored_per_query": r["n_scored"] / sample_q}
            n_cmp = min(sample_q, B)                 # the searcher was created for batches of B
            mine = se.search(queries[:n_cmp], T)
            r = {k2: (v2[:n_cmp] if isinstance(v2, np.ndarray) else v2) for k2, v2 in r.items()}
            cpu["compared_queries"] = n_cmp
            cpu["ids_equal_to_gpu"] = bool(np.array_equal(mine["ids"], r["ids"]))
            cpu["topk_overlap_with_gpu"] = float(np.mean([len(set(a.tolist()) & set(b.tolist())) / max(k, 1)
                                                          for a, b in zip(mine["ids"], r["ids"])]))
            diffs = []
            for qi, (a, b) in enumerate(zip(mine["ids"], r["ids"])):        # scores of items both paths returned
                pa = {int(i): float(s_) for i, s_ in zip(a, mine["scores"][qi])}
                diffs += [abs(pa[int(i)] - float(s_)) for i, s_ in zip(b, r["scores"][qi]) if int(i) in pa]
            cpu["max_abs_score_diff_common_items"] = float(max(diffs)) if diffs else None
            if args.precision == "tensor":            # the bit-exact path, for the record
                sc.set_precision(nb.SCORER_EXACT)
                ex = se.search(queries[:n_cmp], T)
                sc.set_precision(nb.SCORER_TENSOR)
                cpu["exact_path_ids_equal_to_cpu"] = bool(np.array_equal(ex["ids"], r["ids"]))
                cpu["exact_path_scores_bit_equal"] = bool(np.array_equal(ex["scores"].view(np.uint32), r["scores"].view(np.uint32)))
        except Exception as e:
            cpu = {"error": repr(e)[:200]}
    if rank == 0 and world > 1 and args.cpu_shard_sample > 0:
        # the CPU port on ONE shard of the sharded corpus (rank 0's rows, the per-shard beams): a whole query costs the
        # host `world` of these, so queries/s of the CPU port on the full corpus = this / world
        try:
            from oracle import oracle as orc
            cores = os.cpu_count() or 1
            oix = orc.Index(sh["emb"], sh["item_ids"], sh["ep"].astype(np.int32), [v.astype(np.int32) for v in sh["values"]], sh["row_splits"])
            om = orc.Mlp(*sw.mlp_weights(seed=3))
            nq = args.cpu_shard_sample
            r = oix.search_batch_mlp(om, queries[:nq], Ts, nthreads=cores)
            mine = se.search(queries[:nq], Ts)
            extra["cpu_port_one_shard"] = {
                "shard_searches_per_s": nq / r["seconds"], "queries_per_s_full_corpus": nq / r["seconds"] / world, "cores": cores,
                "sample": f"{nq} queries against rank 0's {len(sh['item_ids'])}-row shard with the per-shard beams, one request per core, {r['seconds']:.1f}s",
                "rows_scored_per_shard_search": r["n_scored"] / nq,
                "topk_overlap_with_gpu_shard_search": float(np.mean([len(set(a.tolist()) & set(b.tolist())) / max(Ts[5], 1)
                                                                      for a, b in zip(mine["ids"], r["ids"])]))}
        except Exception as e:
            extra["cpu_port_one_shard"] = {"error": repr(e)[:200]}
    if grp is not None:
        barrier()
        grp.close()
    if rank != 0:
        return None

    pk = peaks()
    qps = B * args.steps / (dev_ms / 1000.0)
    rows = prof["rows_scored"]
    score_ms = prof["ms"]["score"]
    n_score = max(prof["launches"]["score"], 1)
    ach_tf = rows * 2.0 * MAC_PER_ROW / (score_ms / 1000.0) / 1e12 if score_ms > 0 else 0.0
    stage_tot = sum(prof["ms"].values())
    held = extra.get("recall_held", True) if world > 1 else True
    out = {
        "metric": "queries/sec at fixed recall@200" if held else
                  ("queries/sec at the reported recall@200 (no one-GPU operating point exists for this corpus: see recall_*)"
                   if world > 1 and extra.get("recall_target") is None else
                   "queries/sec (recall NOT held at the one-GPU operating point, see recall_*)"),
        "value": qps, "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args, T, world), shard_level_topn=Ts, scorer_precision=args.precision),
        "exchange": exchange,
        "batches_in_flight": P,
        "e2e": {"value": B * args.steps / (e2e_wall_ms / 1000.0), "unit": "queries/s",
                "h2d_bytes_per_step": B * 128 * 4,
                "d2h_bytes_per_step": (B * k * 12 + B * 4 + 10 * B * 4) if world == 1 else B * k * 12 + B * 4,
                "timing": "wall clock around the public API call with pinned host inputs and host outputs" +
                          (f" ({P} host threads, one searcher + stream each)" if P > 1 else ""),
                "device_ms_per_step": e2e_dev_ms / args.steps},
        "latency_ms": latency_ms,
        "gpu_launches": launches,
        "clocks": clk.summary(),
        "roofline": {
            "kernel": "mlp_exact_kernel (fused row gather + 2x512 MLP, fp32 FFMA)" if args.precision == "exact"
                      else "mlp_tc8_kernel (fused row gather + 2x512 MLP, tcgen05 fp16 hi/lo split, cluster-pair neuron split)",
            "bound": "tensor", "achieved": ach_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
            "frac": ach_tf / pk["tf_sustained"], "peak_source": f"bf16 dense sustained, of {pk['source']}",
            "traffic": _ncu_traffic(args.precision),
            "algorithmic_flops_per_row": 2 * MAC_PER_ROW, "rows_per_launch": rows / n_score,
            "issued_mma_tflops": (3.0 * ach_tf * (2 * (128 * 512 + 512 * 512)) / (2.0 * MAC_PER_ROW)) if args.precision == "tensor" else None,
            "note": "tensor precision issues 3 fp16 MMAs per fp32 product (hi/lo split, |dscore| <= 1e-5); the layer-1 user half is hoisted per query",
            "avg_launch_ms": score_ms / n_score,
            "timed_with": "CUDA events around every scorer launch, in a second pass over the same steps (the value pass runs without events or host syncs)",
            "gather_GBps_inside_kernel": rows * ROW_BYTES / (score_ms / 1000.0) / 1e9 if score_ms > 0 else 0.0,
            "hbm_peak_GBps": pk["hbm_gbs"]},
        "roofline_gather": _gather_roofline(),
        "stages_ms_per_step": {k2: v / args.steps for k2, v in prof["ms"].items()},
        "stage_share": {k2: (v / stage_tot if stage_tot else 0) for k2, v in prof["ms"].items()},
        "profile_pass_ms_per_step": prof_ms / args.steps,
        "exchange_exposed_ms_per_step": (dev_ms - prof_ms) / args.steps if world > 1 else 0.0,
        "rows_scored_per_query": rows / (B * args.steps) * world,
        "rows_scored_per_query_per_shard": rows / (B * args.steps),
        "cpu_baseline": cpu,
    }
    if world > 1:
        exp = B * k_s * 12 * (world - 1) / 1024.0
        if nvl0 and nvl1:
            out["nvlink"] = {"tx_kib_rank0": nvl1[0] - nvl0[0], "rx_kib_rank0": nvl1[1] - nvl0[1], "source": nvl1[2],
                             "what": "GPU 0's NVLink data counters, delta over warm-up + timed steps of the device-resident pass",
                             "expected_tx_kib_per_step": exp, "steps_counted": n_steps}
        else:
            out["nvlink"] = {"counters": "not exposed by this driver stack (NVML field values and nvidia-smi nvlink -gt d both N/A)",
                             "expected_tx_kib_per_step": exp,
                             "evidence": "the peers' windows are cudaIpcOpenMemHandle mappings of other GPUs' HBM; the merged results "
                                         "are checked bit for bit in tests/test_shard_group_2gpu.py, which fails unless the peer stores land"}
    out.update(extra)
    return out


def run_b200_dist(args, T, rank, world, local_rank):
    """N > 1, --mode dist: ONE graph (replicated), the embedding table row-sharded, the global batch partitioned; every
    scoring round the candidates travel to the ranks that own their rows and the scores come back (nann_search_distributed).
    Results are bit-identical to the one-GPU search, so recall IS the one-GPU operating point."""
    import torch
    import torch.distributed as dist
    import nann_b200 as nb
    from nann_b200 import scorer_weights as sw
    from nann_b200.distributed import DistGroup
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    k, b = T[5], args.batch
    B = b * world
    if rank == 0:
        get_shard(args.n_items, 1, 0, dev)                   # one graph over the whole corpus, built once, read by everybody
    dist.barrier()
    full = get_shard(args.n_items, 1, 0, dev)
    n = full["emb"].shape[0]
    per = -(-n // world)
    lo, hi = rank * per, min((rank + 1) * per, n)
    ix = nb.Index.from_arrays_sharded(n, full["emb"][lo:hi], lo, full["item_ids"], full["ep"], full["values"], full["row_splits"], device=local_rank)
    sc = nb.Scorer.mlp(*sw.mlp_weights(seed=3), device=local_rank)
    if args.precision == "tensor":
        sc.set_precision(nb.SCORER_TENSOR)
    se = nb.Searcher(ix, sc, b, T)
    grp = DistGroup(se, rank, world)
    ok = torch.ones(1, device=dev)
    try:
        grp.connect_torch()
    except Exception as e:                                   # no peer mappings in this container
        log(f"[bench] rank {rank}: peer windows unavailable ({e!r})")
        ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if ok.item() == 0.0:                                     # every rank takes the same way out
        grp.close()
        return "fallback"
    n_steps = args.warmup + args.steps
    queries = nix().synthetic_queries(full["emb"], B * n_steps, seed=2)      # the same global batches on every rank
    n_eval = max(world, args.eval_queries // world * world)
    eval_q = nix().synthetic_queries(full["emb"], n_eval, seed=5)
    stream = torch.cuda.Stream(device=dev)
    mine = lambda i: slice(i * B + rank * b, i * B + (rank + 1) * b)          # noqa: E731  this rank's queries of global batch i
    q_dev = torch.from_numpy(queries).to(dev)
    q_pin = torch.from_numpy(queries).pin_memory()
    outs = [(torch.empty((b, k), dtype=torch.int64, device=dev), torch.empty((b, k), dtype=torch.float32, device=dev),
             torch.empty((b,), dtype=torch.int32, device=dev)) for _ in range(2)]

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def step_device(i):
        grp.search(q_dev[mine(i)], T, *outs[i & 1], stream=stream)

    def step_profile(i):
        grp.search(q_dev[mine(i)], T, *outs[i & 1], stream=stream, want_stats=True)

    def step_e2e(i):
        return grp.search(q_pin[mine(i)].numpy(), T)

    step_wall = []

    def timed(fn, profile):
        for w in range(args.warmup):
            fn(w)
        barrier()
        se.set_profile(profile)
        l0 = nb.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        del step_wall[:]
        t0 = time.perf_counter()
        timed.window = [t0, t0]
        e0.record(stream)
        for s_ in range(args.steps):
            ts = time.perf_counter()
            fn(args.warmup + s_)
            step_wall.append(time.perf_counter() - ts)
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        timed.window[1] = t0 + wall
        dev_ms = e0.elapsed_time(e1)
        prof = se.profile() if profile else None
        se.set_profile(False)
        t = torch.tensor([dev_ms, wall * 1000.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t[0].item(), t[1].item(), nb.launch_count() - l0, prof

    with ClockSampler(local_rank) as clk:
        dev_ms, _, launches, _ = timed(step_device, False)
        clk.window(*timed.window)
    grp.check()
    prof_ms, _, _, prof = timed(step_profile, True)
    e2e_dev_ms, e2e_wall_ms, _, _ = timed(step_e2e, False)
    lat = sorted(step_wall)
    latency_ms = {"p50": 1000.0 * lat[len(lat) // 2], "max": 1000.0 * lat[-1], "batch": B,
                  "what": "wall time of one public API call per rank (host queries in, host ids+scores out), all ranks in step"}

    # ---- outside the timed region: recall vs brute force, and identity with the one-GPU search
    extra = {}
    try:
        be = n_eval // world
        _, ids_e, st_e = grp.search(eval_q[rank * be:(rank + 1) * be], T)
        g_ids = [torch.empty((be, k), dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(g_ids, torch.from_numpy(ids_e).to(dev))
        got = torch.cat(g_ids, 0).cpu().numpy()
        emb_dev = torch.from_numpy(full["emb"][lo:hi]).to(dev)
        truth = Truth(nb, sc, emb_dev, torch.from_numpy(full["item_ids"][lo:hi]).to(dev), eval_q, k, world, dev)
        del emb_dev
        extra["recall_at_k_vs_bruteforce"] = truth.recall(got)
        extra["recall_queries"] = n_eval
        if rank == 0:
            ix1 = nb.Index.from_arrays(full["emb"], full["item_ids"], full["ep"], full["values"], full["row_splits"], device=local_rank)
            one = nb.Searcher(ix1, sc, n_eval, T).search(eval_q, T)
            extra["recall_target"] = truth.recall(one["ids"])
            extra["recall_target_source"] = "the unsharded index searched on rank 0, same queries, this run"
            extra["ids_bit_identical_to_one_gpu_search"] = bool(np.array_equal(one["ids"], got))
            extra["recall_held"] = bool(extra["recall_at_k_vs_bruteforce"] >= extra["recall_target"] - RECALL_TOL)
            extra["one_gpu_rows_scored_per_query"] = float(one["n_scored"].sum()) / n_eval
            del ix1
    except Exception as e:
        extra["recall_error"] = repr(e)[:200]
    # ---- replica mode (what the reference's session pool does): the whole index on every GPU, no exchange
    if not args.no_replica:
        try:
            ix1 = nb.Index.from_arrays(full["emb"], full["item_ids"], full["ep"], full["values"], full["row_splits"], device=local_rank)
            se1 = nb.Searcher(ix1, sc, b, T)
            o1 = [(torch.empty((b, k), dtype=torch.int64, device=dev), torch.empty((b, k), dtype=torch.float32, device=dev)) for _ in range(2)]

            def step_replica(i):
                se1.search_async(q_dev[mine(i)], T, *o1[i & 1], stream=stream)

            r_ms, _, _, _ = timed(step_replica, False)
            extra["replica_mode"] = {"value": B * args.steps / (r_ms / 1000.0), "unit": "queries/s", "ms_per_step": r_ms / args.steps,
                                     "what": f"whole {n}-row index on every GPU, {b} queries per GPU and step, no exchange (the reference's "
                                             f"replica model, blaze-benchmark/benchmark/core/model.cc:192-234); same results as the one-GPU search"}
            del se1, ix1
        except Exception as e:
            extra["replica_mode"] = {"error": repr(e)[:200]}
    barrier()
    grp.close()
    if rank != 0:
        return None
    pk = peaks()
    rows = prof["rows_scored"]
    score_ms = prof["ms"]["score"]
    n_score = max(prof["launches"]["score"], 1)
    ach_tf = rows * 2.0 * MAC_PER_ROW / (score_ms / 1000.0) / 1e12 if score_ms > 0 else 0.0
    cfg = workload_config(args, T, 1)
    cfg["workload"] = cfg["workload"].replace(f"batch={args.batch} queries", f"global batch={B} queries ({args.batch} per GPU)") + f", {world} GPUs"
    cfg["parallelism"] = (f"distributed scoring x{world}: ONE graph (replicated with the enter points and item ids), the embedding table "
                          f"row-sharded ({per} rows per GPU), the global batch partitioned; per scoring round the candidates go to the ranks "
                          f"that own their rows and the scores come back (stores into IPC-mapped peer windows over NVLink); weak scaling")
    out = {
        "metric": "queries/sec at fixed recall@200", "value": B * args.steps / (dev_ms / 1000.0), "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(cfg, scorer_precision=args.precision, mode="dist"),
        "exchange": "in-library (csrc/lib_dist.inl): bucket / return kernels store candidate ids and scores into the owners' / sources' windows, "
                    "one-warp flag waits on the same stream; 11 flag round trips per step",
        "e2e": {"value": B * args.steps / (e2e_wall_ms / 1000.0), "unit": "queries/s", "h2d_bytes_per_step": B * 128 * 4,
                "d2h_bytes_per_step": B * k * 12 + B * 4,
                "timing": "wall clock around the public API call with pinned host inputs and host outputs (every rank its slice)",
                "device_ms_per_step": e2e_dev_ms / args.steps},
        "latency_ms": latency_ms, "gpu_launches": launches, "clocks": clk.summary(),
        "roofline": {
            "kernel": "mlp_tc8_kernel (fused row gather + 2x512 MLP, tcgen05 fp16 hi/lo split, cluster-pair neuron split)" if args.precision == "tensor"
                      else "mlp_exact_kernel (fused row gather + 2x512 MLP, fp32 FFMA)",
            "bound": "tensor", "achieved": ach_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach_tf / pk["tf_sustained"],
            "peak_source": f"bf16 dense sustained, of {pk['source']}", "traffic": _ncu_traffic(args.precision),
            "algorithmic_flops_per_row": 2 * MAC_PER_ROW, "rows_per_launch": rows / n_score, "avg_launch_ms": score_ms / n_score,
            "note": "in this mode the CUDA events bracket a whole distributed scoring round (bucket + flag wait + scorer kernel + return + "
                    "flag wait + unbucket), so `achieved` understates the scorer kernel by the exchange time; rows = rows of this rank's "
                    "queries (each rank scores as many rows of other ranks' queries as others score of its own)"},
        "roofline_gather": _gather_roofline(),
        "stages_ms_per_step": {k2: v / args.steps for k2, v in prof["ms"].items()},
        "profile_pass_ms_per_step": prof_ms / args.steps,
        "rows_scored_per_query": rows / (b * args.steps),
        "cpu_baseline": None,
    }
    out.update(extra)
    return out


def _ncu_traffic(precision):
    """DRAM bytes of one launch of the dominant kernel, from the committed `ncu --set full` capture (profiles/)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)[precision]["dram_bytes_per_launch"]
    except Exception:
        return None


def _gather_roofline():
    """the standalone embedding-row gather (GatherV2) against the HBM roofline, from the committed ncu capture"""
    try:
        with open(os.path.join(ROOT, "profiles", "gather_roofline.json")) as f:
            return json.load(f)
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-items", type=int, default=1_000_000)
    ap.add_argument("--batch", type=int, default=256, help="queries per GPU and step")
    ap.add_argument("--ef", type=int, default=200, choices=[200, 400])
    ap.add_argument("--precision", default=os.environ.get("NANN_BENCH_PRECISION", "tensor"), choices=["exact", "tensor"])
    ap.add_argument("--eval-queries", type=int, default=128, help="queries of the recall@k measurement / calibration")
    ap.add_argument("--shard-scale", type=float, default=0.0, help="N>1: fix the per-shard beam scale instead of calibrating it")
    ap.add_argument("--shard-scales", default="", help="N>1: five comma-separated per-beam scales (e.g. the ones a smaller corpus calibrated to)")
    ap.add_argument("--cpu-shard-sample", type=int, default=0,
                    help="N>1: rank 0 also times the CPU port on ITS shard with the per-shard beams for this many queries")
    ap.add_argument("--mode", default=os.environ.get("NANN_BENCH_MODE", "dist"), choices=["shard", "dist"],
                    help="N>1: shard = own HNSW per GPU + one exchange of per-shard top-k (calibrated beams); dist = one graph, "
                         "embedding table row-sharded, distributed scoring (bit-identical to the one-GPU search)")
    ap.add_argument("--no-replica", action="store_true")
    ap.add_argument("--inflight", type=int, default=int(os.environ.get("NANN_BENCH_INFLIGHT", "1")),
                    help="N=1: batches in flight, each on its own searcher + stream (the integer stages of one batch run under "
                         "the scorer of another; the reference's analogue is BLAZE_THREADS_NUM concurrent runs of one op)")
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--build-index-only", action="store_true", help="build + cache the unsharded index files and exit (offline tooling)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    T = EF_TOPN[args.ef]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # Only the JSON line may reach stdout: libraries (NCCL's version banner) print there too, so fd 1 is pointed at
    # stderr for the duration of the run and the line is written to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    if args.build_index_only:
        import torch
        get_shard(args.n_items, 1, 0, "cuda:0" if torch.cuda.is_available() else "cpu")
        return
    if args.impl == "reference":
        out = run_reference(args, T, rank, world)
        if out is not None:
            emit(out)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device; there is no CPU fallback (use --impl reference)")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    try:
        # dist needs ONE graph over the whole corpus on every GPU: beyond 16M rows that build would have to be distributed
        # too (not done), so those corpora use the per-shard-HNSW form
        use_dist = world > 1 and args.mode == "dist" and args.n_items <= BIG
        out = run_b200_dist(args, T, rank, world, local_rank) if use_dist else run_b200(args, T, rank, world, local_rank)
        if isinstance(out, str):                             # peer windows could not be mapped: per-shard HNSW + torch transport
            out = run_b200(args, T, rank, world, local_rank)
        if out is not None:
            emit(out)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
