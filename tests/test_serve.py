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
