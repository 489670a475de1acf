"""REST front-end (nann_b200/serve.py): TF-Serving request/response shapes for the `comm_seq`,
`level_topn` -> `top_k` signature (pb_to_saved_model.py:20-46, README.md:196-221) and the dynamic batcher.
The CPU tests drive the HTTP layer with a stand-in backend (no compute); the GPU test serves a real searcher."""
import threading

import numpy as np
import pytest

serve = pytest.importorskip("nann_b200.serve", reason="needs the built library (package import)")
from starlette.testclient import TestClient  # noqa: E402

T = [100, 200, 200, 200, 200, 200]


def fake_backend(calls):
    def run(users, topn):
        calls.append((users.shape[0], tuple(topn)))
        k = topn[5]
        ids = (np.arange(k, dtype=np.int64)[None, :] + np.round(users[:, :1] * 1000).astype(np.int64))
        status = (users[:, 1] < 0).astype(np.int32) * 3          # a negative second feature plays "TopKV2 n < k"
        return dict(ids=ids, scores=np.zeros((users.shape[0], k), np.float32), status=status)
    return run


def test_columnar_and_row_requests():
    calls = []
    app = serve.create_app(fake_backend(calls), user_floats=4, max_batch_size=8, batch_timeout_us=0)
    with TestClient(app) as c:
        assert c.get("/v1/models/nann").json()["model_version_status"][0]["state"] == "AVAILABLE"
        assert c.get("/v1/models/other").status_code == 404
        r = c.post("/v1/models/nann:predict", json={"inputs": {"comm_seq": [[0.005, 1, 0, 0]], "level_topn": T}})
        assert r.status_code == 200
        out = r.json()["outputs"]
        assert len(out) == 1 and out[0][:3] == [5, 6, 7] and len(out[0]) == 200
        r = c.post("/v1/models/nann:predict", json={"instances": [{"comm_seq": [0.001, 1, 0, 0]}, {"comm_seq": [0.002, 1, 0, 0]}],
                                                    "level_topn": T})
        assert [p[0] for p in r.json()["predictions"]] == [1, 2]
        # errors: malformed body, wrong level_topn length, a failing query -> 400 like an InvalidArgument status
        assert c.post("/v1/models/nann:predict", json={"inputs": {"comm_seq": [[0, 1, 0, 0]]}}).status_code == 400
        assert c.post("/v1/models/nann:predict", json={"inputs": {"comm_seq": [[0, 1, 0, 0]], "level_topn": [1, 2]}}).status_code == 400
        r = c.post("/v1/models/nann:predict", json={"inputs": {"comm_seq": [[0, -1, 0, 0]], "level_topn": T}})
        assert r.status_code == 400 and "InvalidArgument" in r.json()["detail"]
    assert calls[0] == (1, tuple(T))


def test_dynamic_batcher_coalesces_same_level_topn():
    calls = []
    gate = threading.Event()

    def slow_backend(users, topn):
        gate.wait(5)                                     # hold the first call so the others queue up behind it
        return fake_backend(calls)(users, topn)

    b = serve.DynamicBatcher(slow_backend, max_batch_size=4, batch_timeout_us=0)
    T2 = [50] + T[1:]
    out = {}

    def client(i, topn):
        out[i] = b.submit(np.full((1, 4), i / 1000.0 + 1e-6, np.float32), topn)

    th = [threading.Thread(target=client, args=(0, T))]
    th[0].start()
    import time
    time.sleep(0.05)                                     # request 0 is now inside the backend
    for i, topn in ((1, T), (2, T), (3, T), (4, T), (5, T), (6, T2), (7, T)):
        th.append(threading.Thread(target=client, args=(i, topn)))
        th[-1].start()
        time.sleep(0.01)                                 # keep the arrival order deterministic
    gate.set()
    for t in th:
        t.join(10)
    b.close()
    # arrival order is kept; groups never mix level_topn and never exceed max_batch_size
    assert [c[0] for c in calls] == [1, 4, 1, 1, 1]
    assert calls[3][1] == tuple(T2)
    for i in range(8):
        assert int(out[i].ids[0, 0]) == i and out[i].status[0] == 0


@pytest.mark.gpu
def test_serving_a_real_searcher(small_world):
    import nann_b200 as nb
    w = small_world
    ix = nb.Index.from_arrays(w["emb"], w["item_ids"], w["ep"], w["values"], w["row_splits"])
    sc = nb.Scorer.mlp(*w["mlp"])
    se = nb.Searcher(ix, sc, 16, w["T"])
    app = serve.create_app(serve.searcher_backend(se), sc.user_floats, max_batch_size=16, batch_timeout_us=0)
    users = w["queries"][:3]
    want = se.search(users, w["T"])
    with TestClient(app) as c:
        r = c.post("/v1/models/nann:predict", json={"inputs": {"comm_seq": users.tolist(), "level_topn": w["T"]}})
        assert r.status_code == 200
        np.testing.assert_array_equal(np.asarray(r.json()["outputs"], np.int64), want["ids"])
        # a level_topn the searcher was not sized for is the caller's error
        assert c.post("/v1/models/nann:predict", json={"inputs": {"comm_seq": users.tolist(), "level_topn": [9999] * 6}}).status_code == 400


def test_grpc_prediction_service_predict():
    """The reference's smoke test (README.md:200-221) over the wire: a PredictRequest with comm_seq (half) and level_topn
    (int32) built with the wire encoder, sent to /tensorflow.serving.PredictionService/Predict."""
    grpc = pytest.importorskip("grpc")
    from nann_b200 import _pbwire as w
    calls = []
    server, port, batcher = serve.create_grpc_server(fake_backend(calls), user_floats=4, address="127.0.0.1:0", max_batch_size=8,
                                                     batch_timeout_us=0)
    server.start()
    try:
        def request(model, comm_seq, topn):
            spec = w.enc_ld(1, model.encode()) + w.enc_ld(3, b"serving_default")
            entries = [("comm_seq", comm_seq), ("level_topn", topn)]
            return w.enc_ld(1, spec) + b"".join(w.enc_ld(2, w.enc_ld(1, k.encode()) + w.enc_ld(2, w.enc_tensor(v))) for k, v in entries)

        ch = grpc.insecure_channel(f"127.0.0.1:{port}")
        predict = ch.unary_unary("/tensorflow.serving.PredictionService/Predict")
        comm_seq = np.asarray([[0.003, 1, 0, 0], [0.007, 1, 0, 0]], np.float16)       # the reference feeds half
        resp = predict(request("nann", comm_seq, np.asarray(T, np.int32)), timeout=10)
        outs, spec_name = {}, None
        for f, wt, v in w.fields(memoryview(resp)):
            if f == 1:
                kv = {f2: v2 for f2, _, v2 in w.fields(v)}
                outs[bytes(kv[1]).decode()] = w.tensor(kv[2])
            elif f == 2:
                spec_name = bytes({f2: v2 for f2, _, v2 in w.fields(v)}[1]).decode()
        assert spec_name == "nann" and set(outs) == {"top_k"}
        top_k = outs["top_k"]
        assert top_k.dtype == np.int64 and top_k.shape == (2, 200) and top_k[:, 0].tolist() == [3, 7]
        assert calls[-1] == (2, tuple(T))
        with pytest.raises(grpc.RpcError) as e:
            predict(request("other", comm_seq, np.asarray(T, np.int32)), timeout=10)
        assert e.value.code() == grpc.StatusCode.NOT_FOUND
        with pytest.raises(grpc.RpcError) as e:
            predict(request("nann", np.asarray([[0, -1, 0, 0]], np.float16), np.asarray(T, np.int32)), timeout=10)
        assert e.value.code() == grpc.StatusCode.INVALID_ARGUMENT
        ch.close()
    finally:
        server.stop(0)
        batcher.close()


def test_serving_cli_refuses_to_rank_with_random_weights():
    """the CLI must not silently serve a randomly initialised scorer"""
    from nann_b200 import serve

    class _NB:
        class Scorer:
            attention = staticmethod(lambda blob: ("attention", len(blob)))
            mlp = staticmethod(lambda *w: ("mlp", len(w)))
    from nann_b200 import scorer_weights as sw
    with pytest.raises(SystemExit):
        serve.load_scorer(_NB, sw, "attention", None, synthetic=False)
    assert serve.load_scorer(_NB, sw, "attention", None, synthetic=True) == ("attention", sw.ATT_BLOB)
    assert serve.load_scorer(_NB, sw, "mlp", None, synthetic=True) == ("mlp", 5)


def test_batcher_sheds_load_like_blaze_xla_op():
    """waiting pool full -> Internal at once (wait_ms == 0); waited too long -> DeadlineExceeded (wait_ms > 0)
    (blaze_xla_kernel.cc:221-258)"""
    import time
    from nann_b200.serve import DynamicBatcher, Overloaded
    gate = threading.Event()

    def slow_backend(users, topn):
        gate.wait(5)
        k = topn[5]
        return dict(ids=np.zeros((len(users), k), np.int64), scores=np.zeros((len(users), k), np.float32), status=np.zeros(len(users), np.int32))

    b = DynamicBatcher(slow_backend, max_batch_size=1, batch_timeout_us=0, max_waiting=2, wait_ms=0)
    res = []

    def client():
        try:
            b.submit(np.zeros((1, 4), np.float32), [1, 1, 1, 1, 1, 1])
            res.append("ok")
        except Overloaded as e:
            res.append((e.code, str(e)))

    th = [threading.Thread(target=client) for _ in range(8)]
    for t in th:
        t.start()
        time.sleep(0.02)                 # the first one is taken by the worker, two more queue up, the rest are refused
    gate.set()
    [t.join() for t in th]
    assert res.count("ok") >= 3 and any(r != "ok" for r in res)
    assert all(r == "ok" or (r[0] == 13 and "waiting pool is full" in r[1]) for r in res)
    b.close()

    gate.clear()
    b = DynamicBatcher(slow_backend, max_batch_size=1, batch_timeout_us=0, wait_ms=30)
    res.clear()
    th = [threading.Thread(target=client) for _ in range(4)]
    [t.start() for t in th]
    time.sleep(0.2)                      # everybody behind the first request has now waited > 30 ms
    gate.set()
    [t.join() for t in th]
    assert res.count("ok") >= 1 and any(r != "ok" for r in res)
    assert all(r == "ok" or (r[0] == 4 and "blaze wait too long" in r[1]) for r in res)
    b.close()
