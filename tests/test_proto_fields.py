"""Wire-level protobuf handling (nann_b200/_pbwire.py, tf_import.py, serve.py) against the field numbers of the
reference's own .proto files.  tests/golden/proto_fields.json is parsed from the reference checkout by
tests/golden/make_proto_fields.py (tensorflow/core/framework/{graph,node_def,attr_value,tensor,tensor_shape,types}.proto,
core/protobuf/tensor_bundle.proto, core/lib/io/format.h, serving/tensorflow_serving/apis/{predict,model}.proto);
the messages below are encoded / decoded by a generic coder driven ONLY by that fixture, never by the package's own
encoders, so a wrong literal in the package shows up as a mismatch."""
import json
import os
import struct

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
F = json.load(open(os.path.join(HERE, "golden", "proto_fields.json")))


def vi(x):
    out = bytearray()
    x &= (1 << 64) - 1
    while True:
        b = x & 0x7F
        x >>= 7
        out.append(b | (0x80 if x else 0))
        if not x:
            return bytes(out)


def ld(msg, field, payload):
    return vi(F[msg][field] << 3 | 2) + vi(len(payload)) + bytes(payload)


def iv(msg, field, value):
    return vi(F[msg][field] << 3 | 0) + vi(value)


def tensor_proto(a, form):
    a = np.asarray(a)
    dt = {np.dtype("float32"): "DT_FLOAT", np.dtype("float16"): "DT_HALF", np.dtype("int32"): "DT_INT32",
          np.dtype("int64"): "DT_INT64", np.dtype("float64"): "DT_DOUBLE"}[a.dtype]
    shape = b"".join(ld("TensorShapeProto", "dim", iv("TensorShapeProto.Dim", "size", int(d))) for d in a.shape)
    head = iv("TensorProto", "dtype", F["DataType"][dt]) + ld("TensorProto", "tensor_shape", shape)
    if form == "content":
        return head + ld("TensorProto", "tensor_content", a.astype(a.dtype.newbyteorder("<")).tobytes())
    if a.dtype == np.float32:
        return head + ld("TensorProto", "float_val", a.astype("<f4").tobytes())
    if a.dtype == np.float64:
        return head + ld("TensorProto", "double_val", a.astype("<f8").tobytes())
    if a.dtype == np.float16:
        return head + ld("TensorProto", "half_val", b"".join(vi(int(x)) for x in a.view(np.uint16).ravel()))
    name = "int_val" if a.dtype == np.int32 else "int64_val"
    return head + ld("TensorProto", name, b"".join(vi(int(x)) for x in a.ravel()))


def walk(buf):
    i = 0
    while i < len(buf):
        key = shift = 0
        while True:
            b = buf[i]; i += 1
            key |= (b & 0x7F) << shift
            if not b & 0x80:
                break
            shift += 7
        f, wt = key >> 3, key & 7
        if wt == 0:
            v = shift = 0
            while True:
                b = buf[i]; i += 1
                v |= (b & 0x7F) << shift
                if not b & 0x80:
                    break
                shift += 7
        elif wt == 2:
            n = shift = 0
            while True:
                b = buf[i]; i += 1
                n |= (b & 0x7F) << shift
                if not b & 0x80:
                    break
                shift += 7
            v = bytes(buf[i:i + n]); i += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, i)[0]; i += 4
        else:
            v = struct.unpack_from("<Q", buf, i)[0]; i += 8
        yield f, wt, v


def test_fixture_is_current_when_the_reference_is_here():
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "serving")):
        pytest.skip("reference checkout not present (the fixture was generated where it is)")
    import subprocess, sys, tempfile, shutil
    tmp = tempfile.mkdtemp()
    try:
        shutil.copy(os.path.join(HERE, "golden", "make_proto_fields.py"), tmp)
        subprocess.run([sys.executable, os.path.join(tmp, "make_proto_fields.py"), ref], check=True, capture_output=True)
        assert json.load(open(os.path.join(tmp, "proto_fields.json"))) == F
    finally:
        shutil.rmtree(tmp)


@pytest.mark.parametrize("form", ["content", "typed"])
def test_graph_def_reader_uses_the_reference_field_numbers(form):
    from nann_b200 import tf_import
    rng = np.random.default_rng(0)
    consts = {"a/kernel": rng.standard_normal((3, 4)).astype(np.float32), "b": np.arange(5, dtype=np.int32),
              "c": rng.standard_normal(6).astype(np.float16), "d": np.array([7, -8], np.int64)}
    g = b""
    for name, a in consts.items():
        attr = ld("NodeDef", "attr", vi(1 << 3 | 2) + vi(len(b"value")) + b"value" +          # map entry: key = 1, value = 2
                  vi(2 << 3 | 2) + vi(len(ld("AttrValue", "tensor", tensor_proto(a, form)))) + ld("AttrValue", "tensor", tensor_proto(a, form)))
        g += ld("GraphDef", "node", ld("NodeDef", "name", name.encode()) + ld("NodeDef", "op", b"Const") + attr)
    g += ld("GraphDef", "node", ld("NodeDef", "name", b"x") + ld("NodeDef", "op", b"Placeholder") + ld("NodeDef", "input", b"y"))
    got = tf_import.read_graph_def_consts(g)
    assert set(got) == set(consts)
    for k, a in consts.items():
        assert got[k].dtype == a.dtype
        np.testing.assert_array_equal(got[k], a)


def test_predict_request_and_response_use_the_reference_field_numbers():
    from nann_b200 import serve
    seq = np.random.default_rng(1).standard_normal((2, 8)).astype(np.float16)
    topn = np.array([10, 20, 20, 20, 20, 20], np.int32)

    def entry(key, a):                                      # map<string, TensorProto>: key = 1, value = 2
        t = tensor_proto(a, "content")
        return vi(1 << 3 | 2) + vi(len(key)) + key.encode() + vi(2 << 3 | 2) + vi(len(t)) + t

    req = ld("PredictRequest", "model_spec", ld("ModelSpec", "name", b"nann") + ld("ModelSpec", "signature_name", b"serving_default")) + \
        ld("PredictRequest", "inputs", entry("comm_seq", seq)) + ld("PredictRequest", "inputs", entry("level_topn", topn)) + \
        ld("PredictRequest", "output_filter", b"top_k")
    name, sig, inputs = serve.parse_predict_request(req)
    assert (name, sig) == ("nann", "serving_default")
    np.testing.assert_array_equal(inputs["comm_seq"], seq)
    np.testing.assert_array_equal(inputs["level_topn"], topn)
    out = np.arange(6, dtype=np.int64).reshape(2, 3)
    resp = serve.encode_predict_response("nann", "serving_default", {"top_k": out})
    seen = {}
    for f, wt, v in walk(resp):
        if f == F["PredictResponse"]["outputs"]:
            kv = dict((f2, v2) for f2, _, v2 in walk(v))
            t = dict((f2, v2) for f2, _, v2 in walk(kv[2]))
            assert t[F["TensorProto"]["dtype"]] == F["DataType"]["DT_INT64"]
            dims = [dict((f4, v4) for f4, _, v4 in walk(v3))[F["TensorShapeProto.Dim"]["size"]]
                    for f3, _, v3 in walk(t[F["TensorProto"]["tensor_shape"]]) if f3 == F["TensorShapeProto"]["dim"]]
            seen[kv[1].decode()] = np.frombuffer(t[F["TensorProto"]["tensor_content"]], "<i8").reshape(dims)
        elif f == F["PredictResponse"]["model_spec"]:
            ms = dict((f2, v2) for f2, _, v2 in walk(v))
            assert ms[F["ModelSpec"]["name"]] == b"nann" and ms[F["ModelSpec"]["signature_name"]] == b"serving_default"
    np.testing.assert_array_equal(seen["top_k"], out)


def test_checkpoint_index_reader_uses_the_reference_format(tmp_path):
    """BundleHeaderProto / BundleEntryProto field numbers and the table constants of core/lib/io/format.h"""
    from nann_b200 import tf_import
    tf_fmt = F["table_format"]
    assert tf_import._TABLE_MAGIC == tf_fmt["magic"]
    a = np.arange(12, dtype=np.float32).reshape(3, 4)
    b = np.array([5, 6, 7], np.int64)
    data, entries = bytearray(), [(b"", iv("BundleHeaderProto", "num_shards", 1))]
    for name, t in (("layer/bias", b), ("layer/kernel", a)):
        dt = F["DataType"]["DT_FLOAT" if t.dtype == np.float32 else "DT_INT64"]
        shape = b"".join(ld("TensorShapeProto", "dim", iv("TensorShapeProto.Dim", "size", int(d))) for d in t.shape)
        raw = t.tobytes()
        e = iv("BundleEntryProto", "dtype", dt) + ld("BundleEntryProto", "shape", shape) + iv("BundleEntryProto", "shard_id", 0) + \
            iv("BundleEntryProto", "offset", len(data)) + iv("BundleEntryProto", "size", len(raw)) + \
            vi(F["BundleEntryProto"]["crc32c"] << 3 | 5) + b"\0\0\0\0"
        entries.append((name.encode(), e))
        data += raw
    prefix = str(tmp_path / "model.ckpt")
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))

    def block(items):                                        # no prefix sharing: every entry is a restart point
        out, restarts = bytearray(), []
        for k, v in items:
            restarts.append(len(out))
            out += vi(0) + vi(len(k)) + vi(len(v)) + k + v
        for r in restarts or [0]:
            out += struct.pack("<I", r)
        return bytes(out + struct.pack("<I", max(len(restarts), 1)))

    trailer = b"\0" * tf_fmt["block_trailer_size"]           # type 0 = no compression, crc unchecked
    f = bytearray()
    blk = block(entries)
    index_items = [(entries[-1][0], vi(0) + vi(len(blk)))]
    f += blk + trailer
    meta = block([])
    meta_handle = vi(len(f)) + vi(len(meta))
    f += meta + trailer
    ix = block(index_items)
    ix_handle = vi(len(f)) + vi(len(ix))
    f += ix + trailer
    foot = meta_handle + ix_handle
    f += foot + b"\0" * (tf_fmt["footer_length"] - 8 - len(foot)) + struct.pack("<Q", tf_fmt["magic"])
    open(prefix + ".index", "wb").write(bytes(f))
    n_shards, ents = tf_import.read_checkpoint_index(prefix)
    assert n_shards == 1 and set(ents) == {"layer/bias", "layer/kernel"}
    assert ents["layer/kernel"]["shape"] == [3, 4] and ents["layer/kernel"]["dtype"] == F["DataType"]["DT_FLOAT"]
    assert ents["layer/bias"]["offset"] == 0 and ents["layer/kernel"]["offset"] == b.nbytes and ents["layer/kernel"]["size"] == a.nbytes
