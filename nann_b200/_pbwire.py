"""Protobuf wire format, just enough for the TensorFlow messages this package touches without TensorFlow:
GraphDef / NodeDef / AttrValue / TensorProto (tf_import.py) and TF-Serving's PredictRequest / PredictResponse
(serve.py).  Decoding yields (field number, wire type, value) triples; encoding builds length-delimited fields."""
import struct

import numpy as np

DT = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 19: np.float16}   # tensorflow/core/framework/types.proto
DT_OF = {np.dtype(np.float32): 1, np.dtype(np.float64): 2, np.dtype(np.int32): 3, np.dtype(np.int64): 9, np.dtype(np.float16): 19}


# ---- decoding
def varint(buf, i):
    x = shift = 0
    while True:
        b = buf[i]
        i += 1
        x |= (b & 0x7F) << shift
        if not b & 0x80:
            return x, i
        shift += 7


def fields(buf):
    """yield (field_number, wire_type, value) for one message; value = int (varint / fixed) or memoryview (bytes)."""
    i, n = 0, len(buf)
    while i < n:
        key, i = varint(buf, i)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, i = varint(buf, i)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, i)[0]
            i += 8
        elif wt == 2:
            ln, i = varint(buf, i)
            v = buf[i:i + ln]
            i += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, i)[0]
            i += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield f, wt, v


def packed_varints(v):
    out, i = [], 0
    while i < len(v):
        x, i = varint(v, i)
        out.append(x)
    return out


def tensor(buf):
    """TensorProto -> ndarray (dtype 1, tensor_shape 2, tensor_content 4, half_val 13, float_val 5, double_val 6,
    int_val 7, int64_val 10); a short *_val list is extended with its last element, as TensorFlow does."""
    dtype, dims, content, vals = 1, [], None, []
    for f, wt, v in fields(buf):
        if f == 1:
            dtype = v
        elif f == 2:
            for f2, _, v2 in fields(v):
                if f2 == 2:                                   # Dim
                    size = 0
                    for f3, _, v3 in fields(v2):
                        if f3 == 1:
                            size = v3 - (1 << 64) if v3 >> 63 else v3
                    dims.append(size)
        elif f == 4:
            content = bytes(v)
        elif f == 5:                                          # float_val (packed or single fixed32)
            vals += list(np.frombuffer(bytes(v), "<f4")) if wt == 2 else [struct.unpack("<f", struct.pack("<I", v))[0]]
        elif f == 6:
            vals += list(np.frombuffer(bytes(v), "<f8")) if wt == 2 else [struct.unpack("<d", struct.pack("<Q", v))[0]]
        elif f in (7, 10, 13):                                # int_val / int64_val / half_val (bit patterns)
            raw = packed_varints(v) if wt == 2 else [v]
            # negative int32 / int64 values travel as 64-bit two's-complement varints
            vals += [x - (1 << 64) if (f != 13 and x >> 63) else x for x in raw]
    if dtype not in DT:
        return None
    np_t = DT[dtype]
    n = int(np.prod(dims)) if dims else 1
    if content is not None and len(content):
        a = np.frombuffer(content, np.dtype(np_t).newbyteorder("<"), count=n)
    else:
        if dtype == 19:
            a = np.asarray(vals, np.uint16).view(np.float16)
        else:
            a = np.asarray(vals, np_t)
        if a.size == 0:
            a = np.zeros(n, np_t)
        elif a.size < n:
            a = np.concatenate([a, np.full(n - a.size, a[-1], np_t)])
    return np.array(a, np_t).reshape(dims)




# ---- encoding
def enc_varint(x):
    out = bytearray()
    x &= (1 << 64) - 1
    while True:
        b = x & 0x7F
        x >>= 7
        out.append(b | (0x80 if x else 0))
        if not x:
            return bytes(out)


def enc_key(field, wire_type):
    return enc_varint(field << 3 | wire_type)


def enc_ld(field, payload):
    """length-delimited field (strings, bytes, sub-messages, packed repeated)"""
    return enc_key(field, 2) + enc_varint(len(payload)) + bytes(payload)


def enc_tensor(a):
    """ndarray -> TensorProto bytes (dtype, tensor_shape, tensor_content)"""
    a = np.asarray(a, order="C")                      # keeps 0-d arrays 0-d
    shape = b"".join(enc_ld(2, enc_key(1, 0) + enc_varint(int(d))) for d in a.shape)
    return enc_key(1, 0) + enc_varint(DT_OF[a.dtype]) + enc_ld(2, shape) + enc_ld(4, a.astype(a.dtype.newbyteorder("<")).tobytes())
