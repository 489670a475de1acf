"""Weight importer for the reference's frozen graph (nann_b200/tf_import.py): the GraphDef is written
here with a minimal protobuf encoder (TensorFlow is not available), using the variable names Model.forward creates
(NANN_impls/nann/model/model_util.py:32-97, model.py:189-233) and the two forms BatchNorm takes after freezing."""
import struct

import numpy as np
import pytest

tfi = pytest.importorskip("nann_b200.tf_import", reason="needs the built library (package import)")
from nann_b200 import scorer_weights as sw  # noqa: E402


def _vi(x):
    out = bytearray()
    x &= (1 << 64) - 1
    while True:
        b = x & 0x7F
        x >>= 7
        out.append(b | (0x80 if x else 0))
        if not x:
            return bytes(out)


def _ld(field, payload):
    return _vi(field << 3 | 2) + _vi(len(payload)) + payload


def _tensor_proto(a, form="content"):
    dt = {np.dtype(np.float32): 1, np.dtype(np.float16): 19, np.dtype(np.int32): 3}[a.dtype]
    shape = b"".join(_ld(2, _vi(1 << 3 | 0) + _vi(d)) for d in a.shape)
    msg = _vi(1 << 3 | 0) + _vi(dt) + _ld(2, shape)
    if form == "content":
        msg += _ld(4, a.astype(a.dtype.newbyteorder("<")).tobytes())
    elif form == "float_val":                       # packed repeated float
        msg += _ld(5, a.astype("<f4").tobytes())
    elif form == "splat":                           # one value, broadcast to the shape (how TF stores constant fills)
        msg += _vi(5 << 3 | 5) + struct.pack("<f", float(a.ravel()[0]))
    elif form == "half_val":
        msg += _ld(13, b"".join(_vi(int(x)) for x in a.view(np.uint16).ravel()))
    return msg


def _node(name, op, tensor=None):
    msg = _ld(1, name.encode()) + _ld(2, op.encode())
    if tensor is not None:
        msg += _ld(5, _ld(1, b"value") + _ld(2, _ld(8, tensor)))
        msg += _ld(5, _ld(1, b"dtype") + _ld(2, _vi(6 << 3 | 0) + _vi(1)))
    return _ld(1, msg)


def _model_consts(rng, prefix=""):
    c = {}
    A = prefix + "nonlinear_attention/"
    for d, (i, o) in (("dense", (64, 128)), ("dense_1", (128, 256)), ("dense_2", (64, 128)), ("dense_3", (128, 256))):
        c[A + d + "/kernel"] = rng.standard_normal((i, o)).astype(np.float32)
        c[A + d + "/bias"] = rng.standard_normal(o).astype(np.float32)
    c[A + "prelu_q"] = rng.standard_normal(128).astype(np.float32)
    c[A + "prelu_k"] = rng.standard_normal(128).astype(np.float32)
    for n, (i, o) in enumerate(((128, 128), (128, 64), (64, 32)), start=1):
        s = f"{prefix}{n}_dnn/"
        c[s + "fc/kernel"] = rng.standard_normal((i, o)).astype(np.float32)
        c[s + "fc/bias"] = rng.standard_normal(o).astype(np.float32)
        c[s + "bn/gamma"] = rng.uniform(0.5, 1.5, o).astype(np.float32)
        c[s + "bn/beta"] = rng.standard_normal(o).astype(np.float32)
        c[s + "bn/moving_mean"] = rng.standard_normal(o).astype(np.float32)
        c[s + "bn/moving_variance"] = rng.uniform(0.5, 1.5, o).astype(np.float32)
        c[s + "prelu"] = np.full(o, 0.25, np.float32)
    c[prefix + "4_dnn/fc/kernel"] = rng.standard_normal((32, 1)).astype(np.float32)
    return c


def _expected_blob(c, prefix=""):
    A = prefix + "nonlinear_attention/"
    parts = []
    for d1, pr, d2 in (("dense", "prelu_q", "dense_1"), ("dense_2", "prelu_k", "dense_3")):
        parts += [c[A + d1 + "/kernel"], c[A + d1 + "/bias"], c[A + pr], c[A + d2 + "/kernel"], c[A + d2 + "/bias"]]
    for n in (1, 2, 3):
        s = f"{prefix}{n}_dnn/"
        sc, sh = sw.fold_bn(c[s + "bn/gamma"], c[s + "bn/beta"], c[s + "bn/moving_mean"], c[s + "bn/moving_variance"])
        parts += [c[s + "fc/kernel"], c[s + "fc/bias"], sc, sh, c[s + "prelu"]]
    parts.append(c[prefix + "4_dnn/fc/kernel"])
    return np.concatenate([p.ravel() for p in parts]).astype(np.float32)


def test_frozen_graph_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    c = _model_consts(rng, prefix="tower/")
    forms = ["content", "float_val"]
    g = _node("comm_seq", "Placeholder")
    for i, (name, a) in enumerate(c.items()):
        form = "splat" if name.endswith("prelu") else forms[i % 2]        # prelu alphas are constant fills (0.25)
        g += _node(name, "Const", _tensor_proto(a, form))
    g += _node("tower/1_dnn/fc/kernel/Adam", "Const", _tensor_proto(np.zeros((128, 128), np.float32)))   # optimizer slot: ignored
    g += _node("tower/1_dnn/fc/MatMul", "MatMul")
    p = tmp_path / "frozen_graph.pb"
    p.write_bytes(g)
    consts = tfi.read_graph_def_consts(str(p))
    assert set(c) <= set(consts) and "comm_seq" not in consts
    for name, a in c.items():
        np.testing.assert_array_equal(consts[name], a)
    blob = tfi.attention_blob_from_frozen_graph(str(p))
    np.testing.assert_array_equal(blob.view(np.uint32), _expected_blob(c, "tower/").view(np.uint32))
    out = tmp_path / "blob.npy"
    tfi.main([str(p), str(out)])
    np.testing.assert_array_equal(np.load(out), blob)


def test_folded_batchnorm_and_half_constants():
    rng = np.random.default_rng(1)
    c = _model_consts(rng)
    want = _expected_blob(c)
    folded = dict(c)
    for n in (1, 2, 3):                                   # what fold_constants leaves behind: y = x * mul + sub
        s = f"{n}_dnn/"
        sc, sh = sw.fold_bn(c[s + "bn/gamma"], c[s + "bn/beta"], c[s + "bn/moving_mean"], c[s + "bn/moving_variance"])
        for k in ("gamma", "beta", "moving_mean", "moving_variance"):
            del folded[s + "bn/" + k]
        folded[s + "bn/batchnorm/mul"], folded[s + "bn/batchnorm/sub"] = sc, sh
    g = b"".join(_node(k, "Const", _tensor_proto(v)) for k, v in folded.items())
    np.testing.assert_array_equal(tfi.attention_blob_from_consts(tfi.read_graph_def_consts(g)).view(np.uint32), want.view(np.uint32))
    h = np.asarray([1.5, -2.25, 0.1], np.float16)         # DT_HALF constants come as bit patterns in half_val
    got = tfi.read_graph_def_consts(_node("h", "Const", _tensor_proto(h, "half_val")))["h"]
    np.testing.assert_array_equal(got, h)
    with pytest.raises(KeyError):
        tfi.attention_blob_from_consts({k: v for k, v in c.items() if not k.startswith("4_dnn")})


@pytest.mark.gpu
def test_imported_blob_scores_like_the_oracle(oracle):
    import nann_b200 as nb
    rng = np.random.default_rng(2)
    c = _model_consts(rng)
    for k in c:
        if k.endswith("kernel"):
            c[k] = (c[k] / np.sqrt(c[k].shape[0])).astype(np.float32)
    g = b"".join(_node(k, "Const", _tensor_proto(v)) for k, v in c.items())
    blob = tfi.attention_blob_from_consts(tfi.read_graph_def_consts(g))
    s, a = nb.Scorer.attention(blob), oracle.Attn(blob)
    user = (0.01 * rng.random((50, 64))).astype(np.float32)
    table = (rng.standard_normal((500, 64)) / 8).astype(np.float32)
    ids = rng.integers(0, 500, 200).astype(np.int32)
    assert np.abs(nb.score_ids(s, user, table, ids) - a.score(user, table, ids)).max() <= 1e-5
