"""Dependency-free writer of ONNX files (hand-encoded protobuf, field numbers from SURVEY
Appendix B.3) with the graph structure of the two buffalo exports.  The files are complete
models: the C++ reader (csrc/onnx_reader.cpp) loads them, and so does an independent ONNX engine
(cv2.dnn.readNetFromONNX, oracle/dnn_engine.py) -- no onnx / ONNX Runtime / real model files."""
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


# --------------------------------------------------------------------------------------------
# Complete, loadable ONNX models (opset 11, typed graph inputs / outputs, full attributes) of the
# two architectures in their RAW form -- Conv -> BatchNormalization -> PRelu / Relu / Sigmoid /
# Resize / Add / Flatten / Gemm, nothing folded -- so that an independent ONNX engine
# (cv2.dnn.readNetFromONNX, see oracle/dnn_engine.py) can execute them.  The raw parameters are
# derived from a canonical (folded) weight dict: for a conv that the canonical form stores as
# conv+bias, the file holds a bias-free Conv and a BatchNormalization with random running
# statistics whose fold reproduces the canonical tensors up to fp32 rounding.

def tensor_i64(name: str, vals) -> bytes:
    arr = np.asarray(vals, np.int64).reshape(-1)
    out = _key(1, 0) + _varint(arr.size)
    out += _key(2, 0) + _varint(7)                      # data_type = INT64
    out += _ld(9, arr.tobytes())
    out += _ld(8, name.encode())
    return out


def tensor_f32_empty(name: str) -> bytes:
    return _key(1, 0) + _varint(0) + _key(2, 0) + _varint(1) + _ld(8, name.encode())


def attr_ints(name: str, vals) -> bytes:
    return _ld(1, name.encode()) + b"".join(_key(8, 0) + _varint(int(v)) for v in vals) + _key(20, 0) + _varint(7)


def attr_s(name: str, v: str) -> bytes:
    return _ld(1, name.encode()) + _ld(4, v.encode()) + _key(20, 0) + _varint(3)


def value_info(name: str, shape) -> bytes:
    dims = b"".join(_ld(1, _key(1, 0) + _varint(int(d))) for d in shape)
    ttype = _key(1, 0) + _varint(1) + _ld(2, dims)      # elem_type FLOAT, shape
    return _ld(1, name.encode()) + _ld(2, _ld(1, ttype))


class FullGraph(GraphBuilder):
    """GraphBuilder + typed inputs / outputs + opset import; node order can be shuffled."""

    def __init__(self, opset: int = 11):
        super().__init__()
        self.inputs, self.outputs, self.opset, self.edges = [], [], opset, []

    def init_i64(self, vals) -> str:
        nm = self.name()
        self.inits.append(tensor_i64(nm, vals))
        return nm

    def init_empty(self) -> str:
        nm = self.name()
        self.inits.append(tensor_f32_empty(nm))
        return nm

    def conv(self, x, wt, b=None, stride=1, group=1):
        k = wt.shape[-1]
        ins = [x, self.init(wt)] + ([self.init(b)] if b is not None else [])
        return self.add("Conv", ins, [attr_ints("dilations", [1, 1]), attr_i("group", group),
                                      attr_ints("kernel_shape", [k, k]), attr_ints("pads", [k // 2] * 4),
                                      attr_ints("strides", [stride, stride])])

    def add_named(self, op, inputs, out_name, attrs=()):
        self.nodes.append(node(op, inputs, [out_name], attrs))
        self.edges.append((list(inputs), out_name))
        return out_name

    def add(self, op, inputs, attrs=()) -> str:
        out = super().add(op, inputs, attrs)
        self.edges.append((list(inputs), out))
        return out

    def _random_topological_order(self, seed):
        """A random VALID node order (ONNX requires producers before consumers; any such order is a
        legal export): Kahn's algorithm picking a random ready node each time."""
        rng = np.random.default_rng(seed)
        produced_by = {out: i for i, (_, out) in enumerate(self.edges)}
        deps = [{produced_by[x] for x in ins if x in produced_by} for ins, _ in self.edges]
        done, order = set(), []
        ready = [i for i, d in enumerate(deps) if not d]
        while ready:
            i = ready.pop(int(rng.integers(0, len(ready))))
            order.append(i)
            done.add(i)
            for j, d in enumerate(deps):
                if j not in done and j not in ready and d <= done:
                    ready.append(j)
        assert len(order) == len(self.nodes)
        return order

    def model(self, shuffle_seed=None, topological: bool = True) -> bytes:
        """shuffle_seed: permute the stored node order -- a random valid topological order by default
        (what a different exporter / optimizer may emit), or with topological=False an arbitrary
        permutation (not valid ONNX; only the product's reader is expected to cope)."""
        nodes = list(self.nodes)
        if shuffle_seed is not None and topological:
            nodes = [self.nodes[i] for i in self._random_topological_order(shuffle_seed)]
        elif shuffle_seed is not None:
            np.random.default_rng(shuffle_seed).shuffle(nodes)
        g = b"".join(_ld(1, n) for n in nodes) + _ld(2, b"g") + b"".join(_ld(5, t) for t in self.inits)
        g += b"".join(_ld(11, v) for v in self.inputs) + b"".join(_ld(12, v) for v in self.outputs)
        opset = _ld(1, b"") + _key(2, 0) + _varint(self.opset)
        return _key(1, 0) + _varint(6) + _ld(2, b"fr-b200-tests") + _ld(7, g) + _ld(8, opset)


class _RawBN:
    """Splits a canonical conv+bias (or affine) into raw parameters with random running statistics."""

    def __init__(self, seed, eps=1e-5):
        self.rng, self.eps = np.random.default_rng(seed), eps

    def stats(self, c):
        var = self.rng.uniform(0.3, 3.0, c)
        mean = self.rng.normal(0, 0.5, c)
        gamma = np.exp(self.rng.normal(0, 0.5, c)) * np.where(self.rng.uniform(size=c) < 0.05, -1.0, 1.0)
        return gamma, mean, var

    def affine(self, scale, shift):
        """BN with gamma / sqrt(var + eps) == scale and beta - mean * that == shift."""
        c = scale.shape[0]
        _, mean, var = self.stats(c)
        gamma = scale.astype(np.float64) * np.sqrt(var + self.eps)
        beta = shift.astype(np.float64) + mean * scale.astype(np.float64)
        return [a.astype(np.float32) for a in (gamma, beta, mean, var)]

    def unfold(self, wt, b):
        """conv(w, b) == BN(conv(w_raw)): returns w_raw and the BN's (gamma, beta, mean, var)."""
        c = wt.shape[0]
        gamma, mean, var = self.stats(c)
        s = gamma / np.sqrt(var + self.eps)
        w_raw = wt.astype(np.float64) / s[:, None, None, None]
        beta = b.astype(np.float64) + mean * s
        return w_raw.astype(np.float32), [a.astype(np.float32) for a in (gamma, beta, mean, var)]


def emit_rec_full(w: dict, path: str, raw: bool = True, seed: int = 0, batch: int = 1, shuffle_seed=None,
                  topological: bool = True):
    """w600k_r50-shaped graph.  raw=True: Conv (no bias) -> BatchNormalization everywhere the
    training graph has one (arcface_torch iresnet: bn after the stem conv, bn2 / bn3 inside every
    IBasicBlock, bn after the downsample conv); raw=False: the form torch.onnx leaves after its
    eval-mode Conv+BN fusion (Conv with bias; only the pre-conv bn1, bn2 and features stay)."""
    from oracle import nets
    g = FullGraph()
    rb = _RawBN(seed)
    eps = rb.eps

    def bn_node(x, params):
        return g.add("BatchNormalization", [x] + [g.init(p) for p in params],
                     [attr_f("epsilon", eps), attr_f("momentum", 0.9)])

    def conv_bn(x, name, stride=1):
        wt, b = w[name + ".w"], w[name + ".b"]
        if not raw:
            return g.conv(x, wt, b, stride)
        w_raw, params = rb.unfold(wt, b)
        return bn_node(g.conv(x, w_raw, None, stride), params)

    g.inputs.append(value_info("input.1", (batch, 3, 112, 112)))
    x = conv_bn("input.1", "stem")
    x = g.add("PRelu", [x, g.init(w["stem.prelu"].reshape(-1, 1, 1))])
    for li, (nb, _) in enumerate(nets.REC_LAYERS):
        for bi in range(nb):
            p = f"l{li}.{bi}"
            y = bn_node(x, rb.affine(w[p + ".bn1.scale"], w[p + ".bn1.shift"]))
            y = conv_bn(y, p + ".conv1")
            y = g.add("PRelu", [y, g.init(w[p + ".prelu"].reshape(-1, 1, 1))])
            y = conv_bn(y, p + ".conv2", 2 if bi == 0 else 1)
            sc = conv_bn(x, p + ".ds", 2) if bi == 0 else x
            x = g.add("Add", [y, sc])
    x = bn_node(x, rb.affine(w["bn2.scale"], w["bn2.shift"]))
    x = g.add("Flatten", [x], [attr_i("axis", 1)])
    x = g.add("Gemm", [x, g.init(w["fc.w"]), g.init(w["fc.b"])],
              [attr_f("alpha", 1.0), attr_f("beta", 1.0), attr_i("transB", 1)])
    params = rb.affine(w["feat.scale"], w["feat.shift"])
    out = g.add_named("BatchNormalization", [x] + [g.init(p) for p in params], "683",
                      [attr_f("epsilon", eps), attr_f("momentum", 0.9)])
    g.outputs.append(value_info(out, (batch, 512)))
    open(path, "wb").write(g.model(shuffle_seed, topological))


DET_OUT_NAMES = ["score_8", "score_16", "score_32", "bbox_8", "bbox_16", "bbox_32", "kps_8", "kps_16", "kps_32"]


def emit_det_full(w: dict, path: str, raw: bool = True, seed: int = 0, bbox_scales=(1.0, 1.0, 1.0), batch: int = 1,
                  size: int = 640, flatten_outputs: bool = True, shuffle_seed=None, topological: bool = True):
    """det_500m-shaped graph (mmdet SCRFD: MobileNetV1-style backbone with Conv-BN-ReLU, PAFPN with
    biased convs and no norm, depthwise-separable head towers with BN, biased 3x3 predictors, Scale
    on the bbox branch, sigmoid on the scores inside the graph).  With flatten_outputs the nine
    outputs are [A, c] like the buffalo export (Transpose -> Reshape); otherwise the predictors' NCHW
    maps are the outputs (for engines whose Reshape support is shaky)."""
    from oracle import nets
    g = FullGraph()
    rb = _RawBN(seed)
    eps = rb.eps

    def conv_bn_relu(x, name, stride=1, group=1):
        wt, b = w[name + ".w"], w[name + ".b"]
        if raw:
            w_raw, params = rb.unfold(wt, b)
            x = g.conv(x, w_raw, None, stride, group)
            x = g.add("BatchNormalization", [x] + [g.init(p) for p in params],
                      [attr_f("epsilon", eps), attr_f("momentum", 0.9)])
        else:
            x = g.conv(x, wt, b, stride, group)
        return g.add("Relu", [x])

    def dwsep(x, name, stride):
        c = w[name + ".dw.w"].shape[0]
        x = conv_bn_relu(x, name + ".dw", stride, c)
        return conv_bn_relu(x, name + ".pw")

    def plain(x, name, stride=1):
        return g.conv(x, w[name + ".w"], w[name + ".b"], stride)

    g.inputs.append(value_info("input.1", (batch, 3, size, size)))
    x = conv_bn_relu("input.1", "stem", 2)
    x = dwsep(x, "b0", 1)
    feats = []
    for si, (nb, _) in enumerate(nets.DET_STAGES):
        for bi in range(nb):
            x = dwsep(x, f"s{si}.{bi}", 2 if bi == 0 else 1)
        if si >= 1:
            feats.append(x)
    lat = [plain(f, f"lat{i}") for i, f in enumerate(feats)]
    for i in (2, 1):
        up = g.add("Resize", [lat[i], g.init_empty(), g.init(np.array([1, 1, 2, 2], np.float32))],
                   [attr_s("coordinate_transformation_mode", "asymmetric"), attr_s("mode", "nearest"),
                    attr_s("nearest_mode", "floor")])
        lat[i - 1] = g.add("Add", [lat[i - 1], up])
    inter = [plain(lat[i], f"fpn{i}") for i in range(3)]
    for i in range(2):
        inter[i + 1] = g.add("Add", [inter[i + 1], plain(inter[i], f"down{i}", 2)])
    outs = [inter[0]] + [plain(inter[i], f"pafpn{i - 1}") for i in (1, 2)]
    results = {}
    for i, f in enumerate(outs):
        t = dwsep(f, f"h{i}.t0", 1)
        t = dwsep(t, f"h{i}.t1", 1)
        hw = (size // nets.DET_STRIDES[i]) ** 2
        for kind, c in (("cls", 1), ("reg", 4), ("kps", 10)):
            name = f"h{i}.{kind}"
            wt, b = w[name + ".w"], w[name + ".b"]
            s = float(bbox_scales[i]) if kind == "reg" else 1.0
            if s != 1.0:
                wt, b = wt / np.float32(s), b / np.float32(s)
            y = g.conv(f if False else t, wt, b)
            if kind == "reg" and s != 1.0:
                y = g.add("Mul", [y, g.init(np.array(s, np.float32))])
            oname = {"cls": "score", "reg": "bbox", "kps": "kps"}[kind] + f"_{nets.DET_STRIDES[i]}"
            if flatten_outputs:
                y = g.add("Transpose", [y], [attr_ints("perm", [0, 2, 3, 1])])
                y = g.add("Reshape", [y, g.init_i64([batch, hw * 2, c] if batch > 1 else [-1, c])])
                if kind == "cls":
                    y = g.add_named("Sigmoid", [y], oname)
                else:
                    y = g.add_named("Identity", [y], oname)
                shape = (batch, hw * 2, c) if batch > 1 else (hw * 2, c)
            else:
                y = g.add_named("Sigmoid" if kind == "cls" else "Identity", [y], oname)
                shape = (batch, 2 * c, size // nets.DET_STRIDES[i], size // nets.DET_STRIDES[i])
            results[oname] = value_info(oname, shape)
    for nm in DET_OUT_NAMES:
        g.outputs.append(results[nm])
    open(path, "wb").write(g.model(shuffle_seed, topological))
