"""Scorer weights from the reference's TensorFlow artefacts, without TensorFlow.

`NANN_impls/nann/delivery/convert_meta.py:361-398` freezes the trained graph into `frozen_graph.pb`: every
variable of `Model.forward` (`nann/model/model.py:189-233`, `nann/model/model_util.py:32-97`) becomes a `Const`
node that keeps the variable's name.  This module reads those nodes straight from the protobuf wire format (a
GraphDef is `repeated NodeDef node = 1`; a Const carries its value in `attr["value"].tensor`) and assembles the
fp32 blob `nann_scorer_create_attention` takes (layout: nann_b200/scorer_weights.py), folding BatchNorm exactly as
the inference graph does (`gamma * rsqrt(moving_variance + 1e-3)`, `beta - moving_mean * scale`).  Both forms are
accepted: the four BatchNorm variables as separate constants (plain `freeze_graph`) or already folded by the
`fold_constants` transform into `bn/batchnorm/mul` and `bn/batchnorm/sub`.

`read_checkpoint` does the same for a V2 training checkpoint (TensorBundle: `<prefix>.index` is a leveldb-format table of
BundleEntryProto, the values live in `<prefix>.data-*`), so `main.py test` weights can be taken from `model_dir` directly.

    python -m nann_b200.tf_import frozen_graph.pb attention_blob.npy
    python -m nann_b200.tf_import model_dir/model.ckpt-12345 attention_blob.npy
"""
import sys

import numpy as np

from . import scorer_weights as sw
from ._pbwire import fields as _fields, tensor as _tensor


def read_graph_def_consts(path_or_bytes):
    """-> {node name: ndarray} for every Const node of a serialized GraphDef (e.g. frozen_graph.pb)."""
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray, memoryview)) else open(path_or_bytes, "rb").read()
    out = {}
    for f, wt, node in _fields(memoryview(data)):
        if f != 1 or wt != 2:
            continue
        name, op, value = None, None, None
        for f2, wt2, v2 in _fields(node):
            if f2 == 1:
                name = bytes(v2).decode()
            elif f2 == 2:
                op = bytes(v2).decode()
            elif f2 == 5:                                     # attr map entry: key = 1, value = 2 (AttrValue)
                key, av = None, None
                for f3, _, v3 in _fields(v2):
                    if f3 == 1:
                        key = bytes(v3).decode()
                    elif f3 == 2:
                        av = v3
                if key == "value" and av is not None:
                    for f4, _, v4 in _fields(av):
                        if f4 == 8:                           # AttrValue.tensor
                            value = v4
        if op == "Const" and name and value is not None:
            t = _tensor(value)
            if t is not None:
                out[name] = t
    return out


# ------------------------------------------------------------------------------------------------
# TensorBundle checkpoints (tf.train.Saver V2: <prefix>.index + <prefix>.data-XXXXX-of-YYYYY)
# ------------------------------------------------------------------------------------------------
# The .index file is a leveldb-format table (tensorflow/core/lib/io/table*): blocks of prefix-compressed
# (key, value) entries + a restart array, each followed by a 5-byte trailer (compression type, masked crc32c);
# a 48-byte footer holds the handles (offset, size varints) of the metaindex and index blocks and the magic
# 0xdb4775248b80fb57.  Keys are tensor names, values BundleEntryProto {dtype 1, shape 2, shard_id 3, offset 4,
# size 5, crc32c 6, slices 7}; the key "" carries BundleHeaderProto {num_shards 1}.
_TABLE_MAGIC = 0xdb4775248b80fb57


def _block_entries(buf, offset, size):
    from ._pbwire import varint
    if buf[offset + size] != 0:
        raise ValueError("compressed table blocks are not supported (checkpoint index files are written uncompressed)")
    blk = memoryview(buf)[offset:offset + size]
    n_restarts = int.from_bytes(blk[size - 4:size], "little")
    end = size - 4 - 4 * n_restarts
    i, key = 0, b""
    while i < end:
        shared, i = varint(blk, i)
        non_shared, i = varint(blk, i)
        vlen, i = varint(blk, i)
        key = key[:shared] + bytes(blk[i:i + non_shared])
        i += non_shared
        yield key, blk[i:i + vlen]
        i += vlen


def read_checkpoint_index(prefix):
    """-> (num_shards, {tensor name: dict(dtype, shape, shard_id, offset, size)}) from <prefix>.index"""
    from ._pbwire import varint
    buf = open(prefix + ".index", "rb").read()
    if len(buf) < 48 or int.from_bytes(buf[-8:], "little") != _TABLE_MAGIC:
        raise ValueError(f"{prefix}.index is not a TensorBundle index (bad table magic)")
    foot = memoryview(buf)[-48:]
    i = 0
    _, i = varint(foot, i); _, i = varint(foot, i)             # metaindex handle
    ix_off, i = varint(foot, i); ix_size, i = varint(foot, i)
    entries, num_shards = {}, 1
    for _, handle in _block_entries(buf, ix_off, ix_size):     # index block: one entry per data block
        d_off, j = varint(handle, 0)
        d_size, j = varint(handle, j)
        for key, val in _block_entries(buf, d_off, d_size):
            if key == b"":
                for f, _, v in _fields(val):
                    if f == 1:
                        num_shards = v
                continue
            e = dict(dtype=1, shape=[], shard_id=0, offset=0, size=0, sliced=False)
            for f, wt, v in _fields(val):
                if f == 1:
                    e["dtype"] = v
                elif f == 2:
                    for f2, _, v2 in _fields(v):
                        if f2 == 2:
                            e["shape"].append(next((v3 for f3, _, v3 in _fields(v2) if f3 == 1), 0))
                elif f == 3:
                    e["shard_id"] = v
                elif f == 4:
                    e["offset"] = v
                elif f == 5:
                    e["size"] = v
                elif f == 7:
                    e["sliced"] = True
            entries[key.decode()] = e
    return num_shards, entries


def read_checkpoint(prefix, names=None):
    """-> {variable name: ndarray} of a V2 checkpoint (e.g. tf.train.latest_checkpoint(model_dir)); numeric dtypes only.
    `names`: optional predicate on the variable name."""
    from ._pbwire import DT
    num_shards, entries = read_checkpoint_index(prefix)
    out, files = {}, {}
    for name, e in entries.items():
        if (names is not None and not names(name)) or e["dtype"] not in DT:
            continue
        if e["sliced"]:
            raise NotImplementedError(f"{name}: partitioned variables (tensor slices) are not supported")
        sid = e["shard_id"]
        if sid not in files:
            files[sid] = np.memmap(f"{prefix}.data-{sid:05d}-of-{num_shards:05d}", np.uint8, "r")
        raw = files[sid][e["offset"]:e["offset"] + e["size"]]
        dt = np.dtype(DT[e["dtype"]]).newbyteorder("<")
        n = int(np.prod(e["shape"])) if e["shape"] else 1
        if raw.size != n * dt.itemsize:
            raise ValueError(f"{name}: {raw.size} bytes in the data file, shape {e['shape']} needs {n * dt.itemsize}")
        out[name] = np.frombuffer(bytes(raw), dt).astype(DT[e["dtype"]]).reshape(e["shape"])
    return out


def attention_blob_from_checkpoint(prefix):
    """the variables of Model.forward straight from a training checkpoint (optimizer slots are ignored)"""
    keep = lambda n: "Adam" not in n and "Momentum" not in n
    return attention_blob_from_consts(read_checkpoint(prefix, keep))


# ------------------------------------------------------------------------------------------------
# variable names of Model.forward -> the attention scorer blob
# ------------------------------------------------------------------------------------------------
def _find(consts, suffix, shape=None):
    hits = [k for k in consts if k == suffix or k.endswith("/" + suffix)]
    hits = [k for k in hits if "/Adam" not in k and "/Momentum" not in k]          # optimizer slots share the prefix
    if shape is not None:
        hits = [k for k in hits if tuple(consts[k].shape) == tuple(shape)]
    if len(hits) != 1:
        raise KeyError(f"expected exactly one constant named */{suffix}{'' if shape is None else ' of shape ' + str(shape)}, found {hits}")
    return np.asarray(consts[hits[0]], np.float32)


def attention_blob_from_consts(consts, bn_eps=1e-3):
    """Blob for nann_scorer_create_attention from the constants of a frozen `Model.forward` graph.
    Names (model_util.py:70-97: nonlinear_attention/dense{,_1,_2,_3} + prelu_q/prelu_k in creation order q, q_, k, k_;
    model_util.py:32-67: {1,2,3}_dnn/{fc,bn,prelu}, 4_dnn/fc without bias)."""
    parts = []
    A = "nonlinear_attention/"
    for d1, pr, d2 in (("dense", "prelu_q", "dense_1"), ("dense_2", "prelu_k", "dense_3")):
        parts += [_find(consts, A + d1 + "/kernel", (64, 128)).ravel(), _find(consts, A + d1 + "/bias", (128,)),
                  _find(consts, A + pr, (128,)),
                  _find(consts, A + d2 + "/kernel", (128, 256)).ravel(), _find(consts, A + d2 + "/bias", (256,))]
    for i, (fi, fo) in enumerate(((128, 128), (128, 64), (64, 32)), start=1):
        s = f"{i}_dnn/"
        parts += [_find(consts, s + "fc/kernel", (fi, fo)).ravel(), _find(consts, s + "fc/bias", (fo,))]
        try:                                                  # folded by fold_constants: y = x * mul + sub
            scale, shift = _find(consts, s + "bn/batchnorm/mul", (fo,)), _find(consts, s + "bn/batchnorm/sub", (fo,))
        except KeyError:
            scale, shift = sw.fold_bn(_find(consts, s + "bn/gamma", (fo,)), _find(consts, s + "bn/beta", (fo,)),
                                      _find(consts, s + "bn/moving_mean", (fo,)), _find(consts, s + "bn/moving_variance", (fo,)),
                                      eps=bn_eps)
        parts += [scale, shift, _find(consts, s + "prelu", (fo,))]
    parts.append(_find(consts, "4_dnn/fc/kernel", (32, 1)).ravel())
    blob = np.concatenate([np.asarray(p, np.float32).ravel() for p in parts]).astype(np.float32)
    assert blob.size == sw.ATT_BLOB, (blob.size, sw.ATT_BLOB)
    return blob


def attention_blob_from_frozen_graph(path):
    return attention_blob_from_consts(read_graph_def_consts(path))


def main(argv=None):
    import os
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 2:
        raise SystemExit(__doc__)
    blob = attention_blob_from_checkpoint(argv[0]) if os.path.exists(argv[0] + ".index") else attention_blob_from_frozen_graph(argv[0])
    np.save(argv[1], blob)
    print(f"{argv[1]}: {blob.size} floats")


if __name__ == "__main__":
    main()
