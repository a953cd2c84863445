"""Minimal ONNX protobuf reader (TEST INFRASTRUCTURE — part of the CPU oracle).

No `onnx` package exists in this image, so the oracle decodes the protobuf wire
format directly.  Field numbers follow onnx.proto:
ModelProto{graph=7, opset_import=8}; GraphProto{node=1, initializer=5, input=11,
output=12, value_info=13}; NodeProto{input=1, output=2, name=3, op_type=4,
attribute=5}; AttributeProto{name=1, f=2, i=3, s=4, t=5, floats=7, ints=8};
TensorProto{dims=1, data_type=2, float_data=4, int64_data=7, name=8, raw_data=9}.

The reference loads the same files through tract / onnxruntime
(crates/zaru/src/nn/mod.rs:259-363); only tests/, bench.py's cpu_baseline and
__graft_entry__.smoke() may import this module.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np


def _varint(buf, pos):
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf):
    """Yield (field_number, wire_type, value) for one message."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fno, wt, v


def _signed(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(wt, v):
    if wt == 0:
        return [_signed(v)]
    out = []
    pos = 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(_signed(x))
    return out


@dataclass
class Node:
    op: str
    name: str
    inputs: list
    outputs: list
    attrs: dict = field(default_factory=dict)


@dataclass
class Graph:
    nodes: list
    initializers: dict            # name -> np.ndarray
    inputs: list                  # [(name, shape)]
    outputs: list                 # [(name, shape)]
    opset: int = 0
    elem_types: dict = field(default_factory=dict)   # graph input/output name -> TensorProto.DataType (1 f32, 10 f16)


_DTYPES = {1: np.float32, 7: np.int64, 10: np.float16, 6: np.int32, 11: np.float64}


def _tensor(buf):
    dims, dtype, raw, name = [], 1, None, ""
    float_data, int64_data = [], []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            dims += _packed_varints(wt, v)
        elif fno == 2:
            dtype = v
        elif fno == 4:
            if wt == 2:
                float_data += list(struct.unpack(f"<{len(v) // 4}f", v))
            else:
                float_data.append(struct.unpack("<f", v)[0])
        elif fno == 7:
            int64_data += _packed_varints(wt, v)
        elif fno == 8:
            name = bytes(v).decode()
        elif fno == 9:
            raw = bytes(v)
    np_dt = _DTYPES[dtype]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np_dt).copy()
    elif float_data:
        arr = np.asarray(float_data, dtype=np_dt)
    elif int64_data:
        arr = np.asarray(int64_data, dtype=np_dt)
    else:
        arr = np.zeros(0, dtype=np_dt)
    return name, arr.reshape(dims) if dims or arr.size == 1 else arr


def _attr(buf):
    name, val = "", None
    floats, ints = [], []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = bytes(v).decode()
        elif fno == 2:
            val = struct.unpack("<f", v)[0]
        elif fno == 3:
            val = _signed(v)
        elif fno == 4:
            val = bytes(v).decode(errors="replace")
        elif fno == 5:
            val = _tensor(v)[1]
        elif fno == 7:
            if wt == 2:
                floats += list(struct.unpack(f"<{len(v) // 4}f", v))
            else:
                floats.append(struct.unpack("<f", v)[0])
        elif fno == 8:
            ints += _packed_varints(wt, v)
    if floats:
        val = floats
    elif ints:
        val = ints
    return name, val


def _node(buf):
    n = Node("", "", [], [])
    for fno, wt, v in _fields(buf):
        if fno == 1:
            n.inputs.append(bytes(v).decode())
        elif fno == 2:
            n.outputs.append(bytes(v).decode())
        elif fno == 3:
            n.name = bytes(v).decode()
        elif fno == 4:
            n.op = bytes(v).decode()
        elif fno == 5:
            k, a = _attr(v)
            n.attrs[k] = a
    return n


def _value_info(buf, types=None):
    name, shape, elem = "", [], 1
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = bytes(v).decode()
        elif fno == 2:
            for f2, _, v2 in _fields(v):
                if f2 == 1:  # tensor_type
                    for f3, _, v3 in _fields(v2):
                        if f3 == 1:  # elem_type
                            elem = v3
                        if f3 == 2:  # shape
                            for f4, _, v4 in _fields(v3):
                                if f4 == 1:  # dim
                                    d = None
                                    for f5, _, v5 in _fields(v4):
                                        if f5 == 1:
                                            d = _signed(v5)
                                    shape.append(d)
    if types is not None:
        types[name] = elem
    return name, shape


def load(path_or_bytes) -> Graph:
    if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
        data = memoryview(bytes(path_or_bytes))
    else:
        with open(path_or_bytes, "rb") as f:
            data = memoryview(f.read())
    graph_buf, opset = None, 0
    for fno, wt, v in _fields(data):
        if fno == 7:
            graph_buf = v
        elif fno == 8:
            dom, ver = "", 0
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    dom = bytes(v2).decode()
                elif f2 == 2:
                    ver = v2
            if dom in ("", "ai.onnx"):
                opset = ver
    g = Graph([], {}, [], [], opset)
    for fno, wt, v in _fields(graph_buf):
        if fno == 1:
            g.nodes.append(_node(v))
        elif fno == 5:
            name, arr = _tensor(v)
            g.initializers[name] = arr
        elif fno == 11:
            g.inputs.append(_value_info(v, g.elem_types))
        elif fno == 12:
            g.outputs.append(_value_info(v, g.elem_types))
    g.inputs = [(n, s) for n, s in g.inputs if n not in g.initializers]
    return g
