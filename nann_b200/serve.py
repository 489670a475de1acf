"""TF-Serving-compatible REST front-end over the batched search, with dynamic batching.

The reference serves `exec.pb` wrapped into a SavedModel whose `serving_default` signature is
`comm_seq`, `level_topn` -> `top_k` (NANN_impls/nann/delivery/pb_to_saved_model.py:20-46) from the
`alinann/nann_serving` TF-Serving image (README.md:196-221: REST on 8501, gRPC on 8500), one `session.run`
per request.  This module keeps that contract on the REST side --

    POST /v1/models/nann:predict
    {"inputs": {"comm_seq": [[...user floats...]], "level_topn": [100, 200, 200, 200, 200, 200]}}
      -> {"outputs": [[item ids ...]]}                    (columnar format; a single named output is returned bare)
    {"instances": [{"comm_seq": [...]} ...], "level_topn": [...]}   is accepted too (row format)
      -> {"predictions": [[item ids ...], ...]}
    GET /v1/models/nann  -> model version status

and on the gRPC side (`create_grpc_server`: `tensorflow.serving.PredictionService/Predict` with inputs `comm_seq`,
`level_topn` and output `top_k`, the call of the reference's smoke test; messages are handled at the protobuf wire
level, no TensorFlow packages needed) -- and coalesces concurrent requests with the same `level_topn` into ONE nann_search_batch call (up to
`max_batch_size` users, waiting at most `batch_timeout_us` for a fuller batch: the knobs blaze-benchmark's
benchmark_conf would carry).  A request whose query fails the way the reference's graph fails (TopKV2 n < k ...)
gets HTTP 400 with the op's message, as TF-Serving reports an InvalidArgument status.

    python -m nann_b200.serve --embs-dir D/embeddings --index-dir D/index --port 8501
"""
import argparse
import os
import queue
import threading
import time

import numpy as np


class _Pending:
    __slots__ = ("users", "topn", "event", "ids", "scores", "status", "error", "t0")

    def __init__(self, users, topn):
        self.users, self.topn = users, tuple(int(t) for t in topn)
        self.event = threading.Event()
        self.ids = self.scores = self.status = self.error = None


class Overloaded(RuntimeError):
    """load shed the way BlazeXlaOp::Schedule sheds it (blaze_xla_kernel.cc:221-258)"""

    def __init__(self, code, message):
        super().__init__(message)
        self.code = code           # 13 Internal ("waiting pool is full" / "blaze wait too long"), 4 DeadlineExceeded


class DynamicBatcher:
    """backend(users f32[B, uf], level_topn) -> dict(ids i64[B,k], scores f32[B,k], status i32[B]);
    requests are served in arrival order, one backend call per group of same-`level_topn` requests.
    Admission control like the reference's BlazeXlaOp: with `wait_ms` == 0 a request is refused at once when `max_waiting`
    requests are queued already (Internal "waiting pool is full", DENSE_MAX_WAITING_COUNT); with `wait_ms` > 0 a request
    that has waited longer than that when its batch is formed is dropped (DeadlineExceeded "blaze wait too long")."""

    def __init__(self, backend, max_batch_size=256, batch_timeout_us=200, max_waiting=0, wait_ms=0):
        self.backend, self.max_batch = backend, int(max_batch_size)
        self.max_waiting, self.wait_s = int(max_waiting), wait_ms * 1e-3
        self.shed = 0
        self.timeout = batch_timeout_us * 1e-6
        self.q = queue.Queue()
        self.batches = []                      # sizes of the backend calls (observability / tests)
        self._stop = False
        self._carry = None
        self.worker = threading.Thread(target=self._run, daemon=True)
        self.worker.start()

    def submit(self, users, topn):
        p = _Pending(np.ascontiguousarray(users, np.float32), topn)
        if p.users.shape[0] > self.max_batch:
            raise ValueError(f"request carries {p.users.shape[0]} users, max_batch_size is {self.max_batch}")
        if self.wait_s == 0 and self.max_waiting > 0 and self.q.qsize() >= self.max_waiting:
            self.shed += 1
            raise Overloaded(13, f"waiting pool is full {self.q.qsize()}")
        p.t0 = time.perf_counter()
        self.q.put(p)
        p.event.wait()
        if p.error is not None:
            raise p.error
        return p

    def close(self):
        self._stop = True
        self.q.put(None)
        self.worker.join(timeout=5)

    def _next(self, timeout):
        if self._carry is not None:
            p, self._carry = self._carry, None
            return p
        try:
            return self.q.get(timeout=timeout) if timeout is not None else self.q.get()
        except queue.Empty:
            return None

    def _run(self):
        while not self._stop:
            first = self._next(None)
            if first is None:
                continue
            group, n = [first], first.users.shape[0]
            deadline = time.perf_counter() + self.timeout
            while n < self.max_batch:
                remaining = deadline - time.perf_counter()
                if remaining <= 0 and self.q.empty() and self._carry is None:
                    break
                p = self._next(max(remaining, 0.0))     # waits at most until the deadline for a fuller batch
                if p is None:
                    break
                if p.topn != first.topn or n + p.users.shape[0] > self.max_batch:
                    self._carry = p            # starts the next group
                    break
                group.append(p)
                n += p.users.shape[0]
            if self.wait_s > 0:                  # requests that waited too long are answered with DeadlineExceeded
                now, live = time.perf_counter(), []
                for g in group:
                    if now - g.t0 > self.wait_s:
                        g.error = Overloaded(4, f"blaze wait too long {int((now - g.t0) * 1e9)}")
                        g.event.set()
                        self.shed += 1
                    else:
                        live.append(g)
                group, n = live, sum(g.users.shape[0] for g in live)
                if not group:
                    continue
            try:
                res = self.backend(np.concatenate([g.users for g in group], 0), list(first.topn))
                self.batches.append(n)
                o = 0
                for g in group:
                    m = g.users.shape[0]
                    g.ids, g.scores, g.status = res["ids"][o:o + m], res["scores"][o:o + m], res["status"][o:o + m]
                    o += m
            except Exception as e:               # the whole group failed (bad level_topn, device error ...)
                for g in group:
                    g.error = e
            for g in group:
                g.event.set()


def create_app(backend, user_floats, model_name="nann", max_batch_size=256, batch_timeout_us=200, max_waiting=0, wait_ms=0):
    from fastapi import FastAPI, HTTPException, Request

    app = FastAPI(title="nann-b200 serving")
    batcher = DynamicBatcher(backend, max_batch_size, batch_timeout_us, max_waiting, wait_ms)
    app.state.batcher = batcher

    @app.get("/v1/models/{name}")
    def status(name: str):
        if name != model_name:
            raise HTTPException(404, f"Could not find any versions of model {name}")
        return {"model_version_status": [{"version": "1", "state": "AVAILABLE", "status": {"error_code": "OK", "error_message": ""}}]}

    @app.post("/v1/models/{name}:predict")
    async def predict(name: str, request: Request):
        if name != model_name:
            raise HTTPException(404, f"Servable not found for request: Latest({name})")
        body = await request.json()
        try:
            if "inputs" in body:                 # columnar
                inp = body["inputs"]
                users, topn, row_format = inp["comm_seq"], inp["level_topn"], False
            else:                                # row format: level_topn is shared by the batch
                users = [i["comm_seq"] for i in body["instances"]]
                topn = body.get("level_topn", body["instances"][0].get("level_topn"))
                row_format = True
            users = np.asarray(users, np.float32).reshape(-1, user_floats)
            topn = [int(t) for t in np.asarray(topn).reshape(-1)]
            if len(topn) != 6:
                raise ValueError("level_topn must have 6 entries")
        except (KeyError, TypeError, ValueError) as e:
            raise HTTPException(400, f"Malformed request: {e}")
        import anyio
        try:
            p = await anyio.to_thread.run_sync(batcher.submit, users, topn)
        except Overloaded as e:                  # TF-Serving maps Internal -> 500, DeadlineExceeded -> 504
            raise HTTPException(504 if e.code == 4 else 500, str(e))
        except Exception as e:                   # NannError / ValueError from the backend call
            raise HTTPException(400, str(e))
        if np.any(p.status != 0):
            bad = int(np.flatnonzero(p.status != 0)[0])
            raise HTTPException(400, f"InvalidArgument: query {bad} failed with status {int(p.status[bad])} "
                                     f"(TopKV2: input must have at least k columns)")
        return {"predictions": p.ids.tolist()} if row_format else {"outputs": p.ids.tolist()}

    return app


# ------------------------------------------------------------------------------------------------
# gRPC: tensorflow.serving.PredictionService/Predict (the call of the reference's smoke test, README.md:200-221)
# ------------------------------------------------------------------------------------------------
# Messages are handled at the protobuf wire level (no tensorflow / tensorflow-serving-api packages here):
#   PredictRequest  { ModelSpec model_spec = 1; map<string, TensorProto> inputs = 2; repeated string output_filter = 3; }
#   PredictResponse { map<string, TensorProto> outputs = 1; ModelSpec model_spec = 2; }
#   ModelSpec       { string name = 1; google.protobuf.Int64Value version = 2; string signature_name = 3; }
def parse_predict_request(data):
    """-> (model name, signature name, {input name: ndarray})"""
    from ._pbwire import fields, tensor
    name, sig, inputs = "", "", {}
    for f, wt, v in fields(memoryview(data)):
        if f == 1 and wt == 2:
            for f2, _, v2 in fields(v):
                if f2 == 1:
                    name = bytes(v2).decode()
                elif f2 == 3:
                    sig = bytes(v2).decode()
        elif f == 2 and wt == 2:                     # map entry: key = 1, value = 2
            key, val = None, None
            for f2, _, v2 in fields(v):
                if f2 == 1:
                    key = bytes(v2).decode()
                elif f2 == 2:
                    val = tensor(v2)
            if key is not None and val is not None:
                inputs[key] = val
    return name, sig, inputs


def encode_predict_response(model_name, signature, outputs):
    from ._pbwire import enc_ld, enc_tensor
    msg = b"".join(enc_ld(1, enc_ld(1, k.encode()) + enc_ld(2, enc_tensor(v))) for k, v in outputs.items())
    return msg + enc_ld(2, enc_ld(1, model_name.encode()) + enc_ld(3, signature.encode()))


def create_grpc_server(backend, user_floats, address="[::]:8500", model_name="nann", max_batch_size=256,
                       batch_timeout_us=200, max_workers=32, max_waiting=0, wait_ms=0):
    """grpc.Server answering /tensorflow.serving.PredictionService/Predict for inputs `comm_seq` (half or float
    [B, user_floats]) and `level_topn` (int32[6]) with output `top_k` (int64 [B, k]), through the same dynamic
    batcher as the REST front-end.  Returns (server, bound port, batcher); call server.start()."""
    from concurrent import futures
    import grpc
    batcher = DynamicBatcher(backend, max_batch_size, batch_timeout_us, max_waiting, wait_ms)

    def predict(request, context):
        try:
            name, sig, inputs = parse_predict_request(request)
        except Exception as e:
            context.abort(grpc.StatusCode.INVALID_ARGUMENT, f"malformed PredictRequest: {e}")
        if name != model_name:
            context.abort(grpc.StatusCode.NOT_FOUND, f"Servable not found for request: Latest({name})")
        if "comm_seq" not in inputs or "level_topn" not in inputs:
            context.abort(grpc.StatusCode.INVALID_ARGUMENT, "input tensors comm_seq and level_topn are required")
        try:
            users = np.asarray(inputs["comm_seq"], np.float32).reshape(-1, user_floats)
            topn = [int(t) for t in np.asarray(inputs["level_topn"]).reshape(-1)]
            if len(topn) != 6:
                raise ValueError("level_topn must have 6 entries")
            p = batcher.submit(users, topn)
        except Overloaded as e:
            context.abort(grpc.StatusCode.DEADLINE_EXCEEDED if e.code == 4 else grpc.StatusCode.INTERNAL, str(e))
        except Exception as e:
            context.abort(grpc.StatusCode.INVALID_ARGUMENT, str(e))
        if np.any(p.status != 0):
            bad = int(np.flatnonzero(p.status != 0)[0])
            context.abort(grpc.StatusCode.INVALID_ARGUMENT, f"query {bad} failed with status {int(p.status[bad])} "
                                                            f"(TopKV2: input must have at least k columns)")
        return encode_predict_response(model_name, sig or "serving_default", {"top_k": np.ascontiguousarray(p.ids, np.int64)})

    handler = grpc.method_handlers_generic_handler(
        "tensorflow.serving.PredictionService",
        {"Predict": grpc.unary_unary_rpc_method_handler(predict)})       # no (de)serializers: raw bytes in and out
    server = grpc.server(futures.ThreadPoolExecutor(max_workers=max_workers))
    server.add_generic_rpc_handlers((handler,))
    port = server.add_insecure_port(address)
    return server, port, batcher


def searcher_backend(searcher):
    """backend over nann_b200.Searcher (one in-flight call at a time: the batcher has a single worker)."""
    def run(users, topn):
        return searcher.search(users, topn)
    return run


def load_scorer(nb, sw, kind, weights, synthetic):
    """the scorer the endpoint ranks with: trained weights from a file, or -- only when asked for explicitly -- seeded
    random ones.  Refuses to start without either: a server ranking by random weights looks healthy and is useless."""
    import numpy as np
    if weights is None:
        if not synthetic:
            raise SystemExit("--scorer-weights is required (an .npy blob, frozen_graph.pb or checkpoint prefix); "
                             "pass --synthetic to serve seeded random weights")
        return nb.Scorer.attention(sw.attention_blob(seed=3)) if kind == "attention" else nb.Scorer.mlp(*sw.mlp_weights(seed=3))
    if kind == "mlp":
        z = np.load(weights)
        return nb.Scorer.mlp(z["W1"], z["b1"], z["W2"], z["b2"], z["w3"])
    from nann_b200 import tf_import
    if weights.endswith(".npy"):
        blob = np.load(weights)
    elif weights.endswith(".pb"):
        blob = tf_import.attention_blob_from_frozen_graph(weights)
    else:
        blob = tf_import.attention_blob_from_checkpoint(weights)
    return nb.Scorer.attention(np.ascontiguousarray(blob, np.float32))


def main():
    import uvicorn
    import nann_b200 as nb
    from nann_b200 import scorer_weights as sw
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--embs-dir", required=True)
    ap.add_argument("--index-dir", required=True)
    ap.add_argument("--max-level-topn", default="100,200,400,400,400,200")
    ap.add_argument("--max-batch-size", type=int, default=256)
    ap.add_argument("--batch-timeout-us", type=int, default=200)
    ap.add_argument("--max-waiting", type=int, default=int(os.environ.get("DENSE_MAX_WAITING_COUNT", "0")),
                    help="refuse requests when this many are queued (BlazeXlaOp's DENSE_MAX_WAITING_COUNT; 0 = never)")
    ap.add_argument("--wait-ms", type=int, default=0, help="drop requests that waited longer (BlazeKernelOptions.wait_ms; 0 = never)")
    ap.add_argument("--scorer", default="attention", choices=["attention", "mlp"],
                    help="attention = the reference's model (Model.forward, 64-d item embeddings); mlp = the 2x512 bench scorer (128-d)")
    ap.add_argument("--scorer-weights", default=None,
                    help="attention: an .npy blob (scorer_weights.py layout), a frozen_graph.pb (convert_meta.py output) or a "
                         "checkpoint prefix; mlp: an .npz with W1,b1,W2,b2,w3")
    ap.add_argument("--synthetic", action="store_true", help="serve seeded random weights (smoke tests only)")
    ap.add_argument("--precision", default=None, choices=["exact", "tensor"], help="mlp only; default tensor")
    ap.add_argument("--host", default="0.0.0.0")
    ap.add_argument("--port", type=int, default=8501)
    ap.add_argument("--grpc-port", type=int, default=8500, help="0 = REST only")
    args = ap.parse_args()
    ix = nb.Index.load(args.embs_dir, args.index_dir)
    sc = load_scorer(nb, sw, args.scorer, args.scorer_weights, args.synthetic)
    if sc.item_dim != ix.dim:
        raise SystemExit(f"--scorer {args.scorer} expects {sc.item_dim}-d item embeddings, the index has {ix.dim}-d rows")
    if args.scorer == "mlp" and (args.precision or "tensor") == "tensor":
        sc.set_precision(nb.SCORER_TENSOR)
    se = nb.Searcher(ix, sc, args.max_batch_size, [int(t) for t in args.max_level_topn.split(",")])
    lock = threading.Lock()                         # one searcher, two front-ends: serialise the backend calls
    backend = searcher_backend(se)

    def locked(users, topn):
        with lock:
            return backend(users, topn)

    if args.grpc_port:
        server, _, _ = create_grpc_server(locked, sc.user_floats, f"{args.host}:{args.grpc_port}", max_batch_size=args.max_batch_size,
                                          batch_timeout_us=args.batch_timeout_us, max_waiting=args.max_waiting, wait_ms=args.wait_ms)
        server.start()
    app = create_app(locked, sc.user_floats, max_batch_size=args.max_batch_size, batch_timeout_us=args.batch_timeout_us,
                     max_waiting=args.max_waiting, wait_ms=args.wait_ms)
    uvicorn.run(app, host=args.host, port=args.port, log_level="warning")


if __name__ == "__main__":
    main()
