"""Closed-loop benchmark harness: the blaze-benchmark role (blaze-benchmark/benchmark/core/benchmark.cc)
on top of nann_executor_run.  Configuration fields keep the reference's names
(blaze-benchmark/benchmark/proto/bench_conf.proto; NANN_impls/nann/benchmark/gen_benchmark_conf.py:16-30:
predictor_num=4, bench_thread_count=4, qps=-1, duration=60); `max_batch_size` is the one addition.

  python -m nann_b200.harness --embs-dir D/embeddings --index-dir D/index --duration 10 --max-batch-size 256
"""
import argparse
import ctypes as C
import json

import numpy as np

from . import _lib
from ._lib import check

HIST_FIELDS = ("count", "min", "max", "mean", "stddev", "median", "p75", "p95", "p98", "p99", "p999")


class BenchConf(C.Structure):
    _fields_ = [("predictor_num", C.c_int), ("bench_thread_count", C.c_int), ("duration_s", C.c_double),
                ("qps", C.c_int), ("max_queue_size", C.c_int), ("max_batch_size", C.c_int),
                ("batch_timeout_us", C.c_int), ("report_interval_s", C.c_int)]


class BenchReport(C.Structure):
    _fields_ = [("seconds", C.c_double), ("throughput_count", C.c_int64), ("mean_rate", C.c_double),
                ("failures", C.c_int64), ("get_predictor_failures", C.c_int64),
                ("latency_us", C.c_double * 11), ("e2e_latency_us", C.c_double * 11), ("batchsize", C.c_double * 11)]


def run_benchmark(index, scorer, level_topn, queries, predictor_num=4, bench_thread_count=4, duration=60.0, qps=-1,
                  max_queue_size=-1, max_batch_size=1, batch_timeout_us=0, report_interval=3, print_reports=False):
    """Returns dict(throughput=..., latency_us={...}, e2e_latency_us={...}, batchsize={...})."""
    L = _lib.lib()
    L.nann_executor_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
    conf = BenchConf(predictor_num, bench_thread_count, float(duration), int(qps), max_queue_size, max_batch_size,
                     batch_timeout_us, report_interval)
    q = np.ascontiguousarray(queries, np.float32).reshape(-1, scorer.user_floats)
    T = (C.c_int32 * 6)(*[int(t) for t in level_topn])
    rep = BenchReport()
    check(L.nann_executor_run(index._h, scorer._h, C.byref(conf), T, C.c_void_p(q.ctypes.data), q.shape[0],
                              int(bool(print_reports)), C.byref(rep)))
    h = lambda a: dict(zip(HIST_FIELDS, [float(x) for x in a]))
    return dict(seconds=rep.seconds, throughput_count=int(rep.throughput_count), throughput=rep.mean_rate,
                failures=int(rep.failures), get_predictor_failures=int(rep.get_predictor_failures),
                latency_us=h(rep.latency_us), e2e_latency_us=h(rep.e2e_latency_us), batchsize=h(rep.batchsize),
                conf=dict(predictor_num=predictor_num, bench_thread_count=bench_thread_count, duration=duration, qps=qps,
                          max_queue_size=max_queue_size, max_batch_size=max_batch_size, batch_timeout_us=batch_timeout_us))


def main():
    from . import Index, Scorer, scorer_weights as sw, index as nix
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--embs-dir", required=True)
    ap.add_argument("--index-dir", required=True)
    ap.add_argument("--level-topn", default="100,200,400,400,400,200")   # gen_runmeta.py:23
    ap.add_argument("--predictor-num", type=int, default=4)
    ap.add_argument("--bench-thread-count", type=int, default=4)
    ap.add_argument("--duration", type=float, default=60)
    ap.add_argument("--qps", type=int, default=-1)
    ap.add_argument("--max-queue-size", type=int, default=-1)
    ap.add_argument("--max-batch-size", type=int, default=1)
    ap.add_argument("--batch-timeout-us", type=int, default=0)
    ap.add_argument("--n-queries", type=int, default=16384)
    ap.add_argument("--precision", default="exact", choices=["exact", "tensor"])
    a = ap.parse_args()
    ix = Index.load(a.embs_dir, a.index_dir)
    sc = Scorer.mlp(*sw.mlp_weights(seed=3))
    if a.precision == "tensor":
        sc.set_precision(_lib.SCORER_TENSOR)
    emb = np.load(f"{a.embs_dir}/item_embs.npy", mmap_mode="r")
    queries = nix.synthetic_queries(np.asarray(emb[:min(len(emb), 1 << 20)]), a.n_queries, seed=2)
    T = [int(x) for x in a.level_topn.split(",")]
    r = run_benchmark(ix, sc, T, queries, a.predictor_num, a.bench_thread_count, a.duration, a.qps, a.max_queue_size,
                      a.max_batch_size, a.batch_timeout_us, print_reports=True)
    print(json.dumps(r))


if __name__ == "__main__":
    main()
