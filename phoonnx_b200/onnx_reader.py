"""Minimal ONNX protobuf reader/writer (no ``onnx`` package: it is absent offline and the
engine must not depend on it).

Only the fields the weight loader needs are decoded:
  ModelProto{ir_version=1, producer_name=2, graph=7, opset_import=8, metadata_props=14}
  GraphProto{node=1, name=2, initializer=5, input=11, output=12}
  NodeProto{input=1, output=2, name=3, op_type=4, attribute=5}
  AttributeProto{name=1, f=2, i=3, s=4, t=5, floats=7, ints=8, type=20}
  TensorProto{dims=1, data_type=2, float_data=4, int64_data=7, name=8, raw_data=9}
  ValueInfoProto{name=1}
  StringStringEntryProto{key=1, value=2}

The file format read here is what ``phoonnx_train/export_onnx.py:318-350`` writes
(``torch.onnx.export`` opset 15 + ``metadata_props``).
"""
from __future__ import annotations

import gzip
import struct
from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Tuple

import numpy as np

# ----------------------------------------------------------------------------- decoding


def _varint(buf: memoryview, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _fields(buf: memoryview) -> Iterator[Tuple[int, int, object]]:
    """Yield (field_number, wire_type, value); length-delimited values are memoryviews."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = bytes(buf[pos:pos + 8])
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = bytes(buf[pos:pos + 4])
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        if pos > n:
            raise ValueError("truncated protobuf message")
        yield fno, wt, v


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(v) -> List[int]:
    out = []
    pos, n = 0, len(v)
    while pos < n:
        x, pos = _varint(v, pos)
        out.append(_signed(x))
    return out


_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 6: np.int32, 7: np.int64, 9: np.bool_,
           10: np.float16, 11: np.float64}


@dataclass
class OnnxNode:
    op_type: str
    name: str
    inputs: List[str]
    outputs: List[str]
    attrs: Dict[str, object] = field(default_factory=dict)


@dataclass
class OnnxModel:
    producer: str
    opset: int
    inputs: List[str]
    outputs: List[str]
    initializers: Dict[str, np.ndarray]
    nodes: List[OnnxNode]
    metadata: Dict[str, str]


def _parse_tensor(buf: memoryview) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dtype = 1
    name = ""
    raw = None
    floats: List[float] = []
    int64s: List[int] = []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            if wt == 0:
                dims.append(_signed(v))
            else:
                dims.extend(_packed_varints(v))
        elif fno == 2:
            dtype = v
        elif fno == 8:
            name = bytes(v).decode("utf-8")
        elif fno == 9:
            raw = v
        elif fno == 4:
            if wt == 2:
                floats.extend(np.frombuffer(v, dtype="<f4").tolist())
            else:
                floats.append(struct.unpack("<f", v)[0])
        elif fno == 7:
            if wt == 2:
                int64s.extend(_packed_varints(v))
            else:
                int64s.append(_signed(v))
    if dtype not in _DTYPES:
        raise ValueError(f"tensor {name!r}: unsupported ONNX data_type {dtype}")
    np_dt = np.dtype(_DTYPES[dtype])
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np_dt.newbyteorder("<")).astype(np_dt, copy=True)
    elif floats:
        arr = np.asarray(floats, dtype=np_dt)
    elif int64s:
        arr = np.asarray(int64s, dtype=np_dt)
    else:
        arr = np.zeros((0,), dtype=np_dt)
    count = int(np.prod(dims)) if dims else 1
    if arr.size != count:
        raise ValueError(f"tensor {name!r}: {arr.size} elements for dims {dims}")
    return name, arr.reshape(dims)


def _parse_attr(buf: memoryview) -> Tuple[str, object]:
    name = ""
    val: object = None
    ints: List[int] = []
    floats: List[float] = []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = bytes(v).decode("utf-8")
        elif fno == 2:
            val = struct.unpack("<f", v)[0]
        elif fno == 3:
            val = _signed(v)
        elif fno == 4:
            val = bytes(v)
        elif fno == 5:
            val = _parse_tensor(v)[1]
        elif fno == 7:
            if wt == 2:
                floats.extend(np.frombuffer(v, dtype="<f4").tolist())
            else:
                floats.append(struct.unpack("<f", v)[0])
        elif fno == 8:
            if wt == 2:
                ints.extend(_packed_varints(v))
            else:
                ints.append(_signed(v))
    if ints:
        val = ints
    elif floats and val is None:
        val = floats
    return name, val


def _parse_node(buf: memoryview, keep_attr_ops) -> OnnxNode:
    ins: List[str] = []
    outs: List[str] = []
    name = op = ""
    attr_bufs = []
    for fno, _wt, v in _fields(buf):
        if fno == 1:
            ins.append(bytes(v).decode("utf-8"))
        elif fno == 2:
            outs.append(bytes(v).decode("utf-8"))
        elif fno == 3:
            name = bytes(v).decode("utf-8")
        elif fno == 4:
            op = bytes(v).decode("utf-8")
        elif fno == 5:
            attr_bufs.append(v)
    attrs = {}
    if keep_attr_ops is None or op in keep_attr_ops:
        for ab in attr_bufs:
            k, val = _parse_attr(ab)
            attrs[k] = val
    return OnnxNode(op, name, ins, outs, attrs)


def _value_info_name(buf: memoryview) -> str:
    for fno, _wt, v in _fields(buf):
        if fno == 1:
            return bytes(v).decode("utf-8")
    return ""


def read_onnx(path: str, keep_attr_ops=("Conv", "ConvTranspose", "LeakyRelu", "Div")) -> OnnxModel:
    """Parse an ONNX file (optionally gzip-compressed: ``*.onnx.gz``)."""
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as f:
        data = f.read()
    return parse_onnx_bytes(data, keep_attr_ops)


def parse_onnx_bytes(data: bytes, keep_attr_ops=("Conv", "ConvTranspose", "LeakyRelu", "Div")) -> OnnxModel:
    buf = memoryview(data)
    producer = ""
    opset = 0
    graph = None
    meta: Dict[str, str] = {}
    for fno, _wt, v in _fields(buf):
        if fno == 2:
            producer = bytes(v).decode("utf-8")
        elif fno == 7:
            graph = v
        elif fno == 8:
            for f2, _w2, v2 in _fields(v):
                if f2 == 2:
                    opset = max(opset, int(v2))
        elif fno == 14:
            k = val = ""
            for f2, _w2, v2 in _fields(v):
                if f2 == 1:
                    k = bytes(v2).decode("utf-8")
                elif f2 == 2:
                    val = bytes(v2).decode("utf-8")
            meta[k] = val
    if graph is None:
        raise ValueError("not an ONNX ModelProto: no graph")
    inits: Dict[str, np.ndarray] = {}
    nodes: List[OnnxNode] = []
    g_in: List[str] = []
    g_out: List[str] = []
    for fno, _wt, v in _fields(graph):
        if fno == 1:
            nodes.append(_parse_node(v, keep_attr_ops))
        elif fno == 5:
            name, arr = _parse_tensor(v)
            inits[name] = arr
        elif fno == 11:
            g_in.append(_value_info_name(v))
        elif fno == 12:
            g_out.append(_value_info_name(v))
    g_in = [n for n in g_in if n not in inits]
    return OnnxModel(producer, opset, g_in, g_out, inits, nodes, meta)


# ----------------------------------------------------------------------------- encoding
# (used by phoonnx_b200.modelgen to write exporter-format files for synthetic models)


def _enc_varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _enc_key(fno: int, wt: int) -> bytes:
    return _enc_varint((fno << 3) | wt)


def _enc_bytes(fno: int, payload: bytes) -> bytes:
    return _enc_key(fno, 2) + _enc_varint(len(payload)) + payload


def _enc_str(fno: int, s: str) -> bytes:
    return _enc_bytes(fno, s.encode("utf-8"))


def _enc_int(fno: int, v: int) -> bytes:
    return _enc_key(fno, 0) + _enc_varint(int(v))


_ONNX_DT = {np.dtype(np.float32): 1, np.dtype(np.int64): 7}


def encode_tensor(name: str, arr: np.ndarray) -> bytes:
    arr = np.ascontiguousarray(arr)
    out = bytearray()
    for d in arr.shape:
        out += _enc_int(1, d)
    out += _enc_int(2, _ONNX_DT[arr.dtype])
    out += _enc_str(8, name)
    out += _enc_bytes(9, arr.astype(arr.dtype.newbyteorder("<"), copy=False).tobytes())
    return bytes(out)


def encode_attr(name: str, val) -> bytes:
    out = bytearray(_enc_str(1, name))
    if isinstance(val, float):
        out += _enc_key(2, 5) + struct.pack("<f", val)
        out += _enc_int(20, 1)
    elif isinstance(val, int):
        out += _enc_int(3, val)
        out += _enc_int(20, 2)
    elif isinstance(val, (list, tuple)):
        for x in val:
            out += _enc_int(8, int(x))
        out += _enc_int(20, 7)
    else:
        raise TypeError(type(val))
    return bytes(out)


def encode_node(op_type: str, name: str, inputs, outputs, attrs=None) -> bytes:
    out = bytearray()
    for s in inputs:
        out += _enc_str(1, s)
    for s in outputs:
        out += _enc_str(2, s)
    out += _enc_str(3, name)
    out += _enc_str(4, op_type)
    for k, v in (attrs or {}).items():
        out += _enc_bytes(5, encode_attr(k, v))
    return bytes(out)


def encode_model(initializers: Dict[str, np.ndarray], nodes: List[bytes], inputs: List[str],
                 outputs: List[str], metadata: Dict[str, str], producer: str = "pytorch",
                 opset: int = 15) -> bytes:
    g = bytearray()
    for nb in nodes:
        g += _enc_bytes(1, nb)
    g += _enc_str(2, "main_graph")
    for name, arr in initializers.items():
        g += _enc_bytes(5, encode_tensor(name, arr))
    for n in inputs:
        g += _enc_bytes(11, _enc_str(1, n))
    for n in outputs:
        g += _enc_bytes(12, _enc_str(1, n))
    m = bytearray()
    m += _enc_int(1, 8)
    m += _enc_str(2, producer)
    m += _enc_bytes(7, bytes(g))
    m += _enc_bytes(8, _enc_str(1, "") + _enc_int(2, opset))
    for k, v in metadata.items():
        m += _enc_bytes(14, _enc_str(1, k) + _enc_str(2, str(v)))
    return bytes(m)
