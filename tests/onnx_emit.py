"""Dependency-free writer of minimal ONNX files (hand-encoded protobuf, field numbers from
SURVEY Appendix B.3) with the graph structure of the two buffalo exports, used to test the
C++ initializer reader (csrc/onnx_reader.cpp) without onnx / ONNX Runtime / real model files."""
from __future__ import annotations

import struct

import numpy as np


def _varint(n: int) -> bytes:
    out = bytearray()
    n &= (1 << 64) - 1
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _key(field: int, wire: int) -> bytes:
    return _varint((field << 3) | wire)


def _ld(field: int, payload: bytes) -> bytes:
    return _key(field, 2) + _varint(len(payload)) + payload


def tensor(name: str, arr: np.ndarray, raw: bool = True) -> bytes:
    arr = np.ascontiguousarray(arr, np.float32)
    out = b"".join(_key(1, 0) + _varint(d) for d in arr.shape)
    out += _key(2, 0) + _varint(1)                      # data_type = FLOAT
    if raw:
        out += _ld(9, arr.tobytes())                    # raw_data
    else:
        out += _ld(4, arr.tobytes())                    # packed float_data
    out += _ld(8, name.encode())
    return out


def attr_f(name: str, v: float) -> bytes:
    return _ld(1, name.encode()) + _key(2, 5) + struct.pack("<f", v) + _key(20, 0) + _varint(1)


def attr_i(name: str, v: int) -> bytes:
    return _ld(1, name.encode()) + _key(3, 0) + _varint(v) + _key(20, 0) + _varint(2)


def node(op: str, inputs, outputs, attrs=()) -> bytes:
    out = b"".join(_ld(1, i.encode()) for i in inputs)
    out += b"".join(_ld(2, o.encode()) for o in outputs)
    out += _ld(4, op.encode())
    out += b"".join(_ld(5, a) for a in attrs)
    return out


class GraphBuilder:
    def __init__(self):
        self.nodes, self.inits, self.n = [], [], 0

    def name(self) -> str:
        self.n += 1
        return str(self.n)            # numeric names, like the buffalo exports

    def init(self, arr, raw=True) -> str:
        nm = self.name()
        self.inits.append(tensor(nm, arr, raw))
        return nm

    def add(self, op, inputs, attrs=()) -> str:
        out = self.name()
        self.nodes.append(node(op, inputs, [out], attrs))
        return out

    def model(self) -> bytes:
        g = b"".join(_ld(1, n) for n in self.nodes) + _ld(2, b"g") + b"".join(_ld(5, t) for t in self.inits)
        return _key(1, 0) + _varint(7) + _ld(7, g)       # ir_version, graph


def _bn_raw(rng, scale, shift, eps):
    var = rng.uniform(0.5, 2.0, scale.shape).astype(np.float32)
    mean = rng.normal(0, 0.5, scale.shape).astype(np.float32)
    gamma = (scale * np.sqrt(var + np.float32(eps))).astype(np.float32)
    beta = (shift + mean * scale).astype(np.float32)
    return gamma, beta, mean, var


def emit_rec(w: dict, path: str, seed: int = 0, trans_b: bool = True):
    """IResNet-50 export: Conv(+bias) / PRelu / BatchNormalization / Flatten / Gemm / BatchNormalization."""
    from oracle import nets
    rng = np.random.default_rng(seed)
    g = GraphBuilder()
    eps = 1e-5

    def bn(x, scale, shift):
        ga, be, mu, va = _bn_raw(rng, scale, shift, eps)
        return g.add("BatchNormalization", [x, g.init(ga), g.init(be), g.init(mu), g.init(va)], [attr_f("epsilon", eps)])

    x = g.add("Conv", ["input.1", g.init(w["stem.w"]), g.init(w["stem.b"])])
    x = g.add("PRelu", [x, g.init(w["stem.prelu"].reshape(-1, 1, 1))])
    for li, (nb, _) in enumerate(nets.REC_LAYERS):
        for bi in range(nb):
            p = f"l{li}.{bi}"
            y = bn(x, w[p + ".bn1.scale"], w[p + ".bn1.shift"])
            y = g.add("Conv", [y, g.init(w[p + ".conv1.w"]), g.init(w[p + ".conv1.b"])])
            y = g.add("PRelu", [y, g.init(w[p + ".prelu"].reshape(-1, 1, 1))])
            y = g.add("Conv", [y, g.init(w[p + ".conv2.w"], raw=(bi % 2 == 0)), g.init(w[p + ".conv2.b"])])
            sc = x
            if bi == 0:
                sc = g.add("Conv", [x, g.init(w[p + ".ds.w"]), g.init(w[p + ".ds.b"])])
            x = g.add("Add", [y, sc])
    x = bn(x, w["bn2.scale"], w["bn2.shift"])
    x = g.add("Flatten", [x])
    fcw = w["fc.w"] if trans_b else np.ascontiguousarray(w["fc.w"].T)
    x = g.add("Gemm", [x, g.init(fcw), g.init(w["fc.b"])], [attr_i("transB", 1 if trans_b else 0)])
    bn(x, w["feat.scale"], w["feat.shift"])
    open(path, "wb").write(g.model())


def emit_det(w: dict, path: str, bbox_scales=(1.0, 1.0, 1.0), with_scale_nodes: bool = True):
    """SCRFD export: one Conv(+bias) per canonical conv in execution order (Relu/Sigmoid/Add/
    Resize nodes carry no initializers and are emitted as placeholders), Mul by a scalar after
    each bbox conv (mmdet Scale)."""
    from oracle import nets
    g = GraphBuilder()
    x = "input.1"
    names = [n[:-2] for n, _ in nets.det_tensor_specs() if n.endswith(".w")]
    for nm in names:
        wt, b = w[nm + ".w"], w[nm + ".b"]
        s = 1.0
        if nm.endswith(".reg") and with_scale_nodes:
            s = float(bbox_scales[int(nm[1])])
            wt, b = wt / np.float32(s), b / np.float32(s)
        x = g.add("Conv", [x, g.init(wt), g.init(b)])
        if nm.endswith(".reg") and with_scale_nodes:
            x = g.add("Mul", [x, g.init(np.array(s, np.float32))])
        elif nm.endswith(".cls"):
            x = g.add("Sigmoid", [x])
        else:
            x = g.add("Relu", [x])
    open(path, "wb").write(g.model())
