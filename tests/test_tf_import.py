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


def _write_checkpoint(prefix, tensors, block_size=6):
    """A V2 checkpoint written by hand: data file + leveldb-format index table (several data blocks, prefix-compressed
    keys with a restart point every 2 entries, 5-byte block trailers, 48-byte footer)."""
    from nann_b200._pbwire import enc_varint, enc_key, enc_ld, DT_OF
    data, entries = bytearray(), [(b"", enc_key(1, 0) + enc_varint(1) + enc_ld(3, enc_key(1, 0) + enc_varint(1)))]   # header: num_shards 1
    for name in sorted(tensors):
        a = np.asarray(tensors[name], order="C")          # (ascontiguousarray would turn a scalar into shape (1,))
        raw = a.astype(a.dtype.newbyteorder("<")).tobytes()
        shape = b"".join(enc_ld(2, enc_key(1, 0) + enc_varint(int(d))) for d in a.shape)
        e = enc_key(1, 0) + enc_varint(DT_OF[a.dtype]) + enc_ld(2, shape) + enc_key(4, 0) + enc_varint(len(data)) + \
            enc_key(5, 0) + enc_varint(len(raw)) + enc_key(6, 5) + b"\0\0\0\0"
        entries.append((name.encode(), e))
        data += raw
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))

    def block(items):
        out, restarts, prev = bytearray(), [], b""
        for i, (k, v) in enumerate(items):
            shared = 0
            if i % 2 == 0:
                restarts.append(len(out))
            else:
                while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                    shared += 1
            out += enc_varint(shared) + enc_varint(len(k) - shared) + enc_varint(len(v)) + k[shared:] + v
            prev = k
        for r in restarts or [0]:
            out += struct.pack("<I", r)
        out += struct.pack("<I", max(len(restarts), 1))
        return bytes(out)

    f, index_items = bytearray(), []
    for b0 in range(0, len(entries), block_size):
        items = entries[b0:b0 + block_size]
        blk = block(items)
        index_items.append((items[-1][0], enc_varint(len(f)) + enc_varint(len(blk))))
        f += blk + b"\0" + b"\0\0\0\0"                     # trailer: no compression + (unchecked) crc
    meta = block([])
    meta_handle = enc_varint(len(f)) + enc_varint(len(meta))
    f += meta + b"\0" + b"\0\0\0\0"
    ix = block(index_items)
    ix_handle = enc_varint(len(f)) + enc_varint(len(ix))
    f += ix + b"\0" + b"\0\0\0\0"
    foot = meta_handle + ix_handle
    f += foot + b"\0" * (40 - len(foot)) + struct.pack("<Q", 0xdb4775248b80fb57)
    open(prefix + ".index", "wb").write(bytes(f))


def test_checkpoint_round_trip(tmp_path):
    rng = np.random.default_rng(3)
    c = _model_consts(rng, prefix="")
    extra = {"global_step": np.asarray(12345, np.int64), "1_dnn/fc/kernel/Adam": np.zeros((128, 128), np.float32),
             "1_dnn/fc/kernel/Adam_1": np.ones((128, 128), np.float32), "beta1_power": np.asarray(0.9, np.float32)}
    prefix = str(tmp_path / "model.ckpt-12345")
    _write_checkpoint(prefix, {**c, **extra})
    shards, entries = tfi.read_checkpoint_index(prefix)
    assert shards == 1 and set(entries) == set(c) | set(extra)
    got = tfi.read_checkpoint(prefix)
    for k, v in {**c, **extra}.items():
        np.testing.assert_array_equal(got[k], v)
        assert got[k].dtype == v.dtype and got[k].shape == v.shape
    blob = tfi.attention_blob_from_checkpoint(prefix)
    np.testing.assert_array_equal(blob.view(np.uint32), _expected_blob(c).view(np.uint32))
    out = tmp_path / "blob.npy"
    tfi.main([prefix, str(out)])                           # the CLI recognises a checkpoint prefix by its .index file
    np.testing.assert_array_equal(np.load(out), blob)
    (tmp_path / "bad.index").write_bytes(b"x" * 64)
    with pytest.raises(ValueError):
        tfi.read_checkpoint_index(str(tmp_path / "bad"))
