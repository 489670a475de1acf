"""Scorer weights from the reference's TensorFlow artefacts, without TensorFlow.

`NANN_impls/nann/delivery/convert_meta.py:361-398` freezes the trained graph into `frozen_graph.pb`: every
variable of `Model.forward` (`nann/model/model.py:189-233`, `nann/model/model_util.py:32-97`) becomes a `Const`
node that keeps the variable's name.  This module reads those nodes straight from the protobuf wire format (a
GraphDef is `repeated NodeDef node = 1`; a Const carries its value in `attr["value"].tensor`) and assembles the
fp32 blob `nann_scorer_create_attention` takes (layout: nann_b200/scorer_weights.py), folding BatchNorm exactly as
the inference graph does (`gamma * rsqrt(moving_variance + 1e-3)`, `beta - moving_mean * scale`).  Both forms are
accepted: the four BatchNorm variables as separate constants (plain `freeze_graph`) or already folded by the
`fold_constants` transform into `bn/batchnorm/mul` and `bn/batchnorm/sub`.

    python -m nann_b200.tf_import frozen_graph.pb attention_blob.npy
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
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 2:
        raise SystemExit(__doc__)
    blob = attention_blob_from_frozen_graph(argv[0])
    np.save(argv[1], blob)
    print(f"{argv[1]}: {blob.size} floats")


if __name__ == "__main__":
    main()
