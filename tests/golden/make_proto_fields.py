#!/usr/bin/env python
"""Writes tests/golden/proto_fields.json: the field numbers / enum values of the TensorFlow and TF-Serving messages
this package reads and writes at the wire level (nann_b200/_pbwire.py, tf_import.py, serve.py), parsed from the
.proto files of the reference checkout.  Run where /root/reference exists; the fixture travels, the reference does not."""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
FILES = {
    "tensorflow/tensorflow/core/framework/graph.proto": ["GraphDef"],
    "tensorflow/tensorflow/core/framework/node_def.proto": ["NodeDef"],
    "tensorflow/tensorflow/core/framework/attr_value.proto": ["AttrValue"],
    "tensorflow/tensorflow/core/framework/tensor.proto": ["TensorProto"],
    "tensorflow/tensorflow/core/framework/tensor_shape.proto": ["TensorShapeProto", "TensorShapeProto.Dim"],
    "tensorflow/tensorflow/core/framework/types.proto": ["enum DataType"],
    "tensorflow/tensorflow/core/protobuf/tensor_bundle.proto": ["BundleHeaderProto", "BundleEntryProto"],
    "serving/tensorflow_serving/apis/predict.proto": ["PredictRequest", "PredictResponse"],
    "serving/tensorflow_serving/apis/model.proto": ["ModelSpec"],
}


def strip_comments(s):
    s = re.sub(r"//[^\n]*", "", s)
    return re.sub(r"/\*.*?\*/", "", s, flags=re.S)


def block(s, start):
    """text between the braces that open at or after `start`"""
    i = s.index("{", start)
    depth, j = 1, i + 1
    while depth:
        depth += {"{": 1, "}": -1}.get(s[j], 0)
        j += 1
    return s[i + 1:j - 1]


def fields_of(body):
    """top-level fields of a message body (nested message / enum / oneof braces are flattened for oneof only)"""
    out = {}
    flat, depth, i = "", 0, 0
    # keep oneof bodies (their fields belong to the message), drop nested message / enum bodies
    while i < len(body):
        m = re.compile(r"\b(message|enum|oneof)\s+\w+\s*\{").match(body, i)
        if m:
            inner = block(body, m.start())
            if m.group(1) == "oneof":
                flat += inner
            i = body.index("{", m.start()) + len(inner) + 2
            continue
        flat += body[i]
        i += 1
    for m in re.finditer(r"(?:repeated\s+|optional\s+)?(map\s*<[^>]+>|[\w.]+)\s+(\w+)\s*=\s*(\d+)", flat):
        out[m.group(2)] = int(m.group(3))
    return out


out = {"_source": "parsed from the reference checkout by tests/golden/make_proto_fields.py"}
for rel, names in FILES.items():
    src = strip_comments(open(os.path.join(REF, rel)).read())
    for name in names:
        if name.startswith("enum "):
            body = block(src, re.search(r"\benum\s+%s\b" % name[5:], src).start())
            out[name[5:]] = {m.group(1): int(m.group(2)) for m in re.finditer(r"(\w+)\s*=\s*(\d+)\s*;", body)}
            continue
        body = src
        for part in name.split("."):
            body = block(body, re.search(r"\bmessage\s+%s\b" % part, body).start())
        out[name] = fields_of(body)
        out[name]["_file"] = rel
# leveldb-style table constants of the checkpoint index file (tensorflow/core/lib/io/format.h)
fmt = open(os.path.join(REF, "tensorflow/tensorflow/core/lib/io/format.h")).read()
max_handle = int(re.search(r"kMaxEncodedLength\s*=\s*(\d+)\s*\+\s*(\d+)", fmt).group(1)) * 2
out["table_format"] = {
    "_file": "tensorflow/tensorflow/core/lib/io/format.h",
    "magic": int(re.search(r"kTableMagicNumber\s*=\s*(0x[0-9a-fA-F]+)", fmt).group(1), 16),
    "block_trailer_size": int(re.search(r"kBlockTrailerSize\s*=\s*(\d+)", fmt).group(1)),
    "footer_length": 2 * max_handle + 8,            # Footer::kEncodedLength = 2 * BlockHandle::kMaxEncodedLength + 8
}
with open(os.path.join(HERE, "proto_fields.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print("wrote proto_fields.json:", {k: len(v) for k, v in out.items() if isinstance(v, dict)})
