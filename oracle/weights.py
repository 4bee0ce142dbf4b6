"""Weight generators for the oracle side.

TEST INFRASTRUCTURE (see oracle/__init__.py): nothing in the product imports this.

The reference reads its weights from models/*.onnx (src/face_detector.cpp:20-90,
src/face_recognizer.cpp:21-91); the files are not shipped (models/README.md:21-31) and there is no
network, so every test / bench run needs synthetic weights of the two architectures.  Two kinds:

* ``seeded(model, seed)`` -- a pure-NumPy twin of the product's seeded random-init
  (csrc/weights.cpp, splitmix64 per element).  It is bit-identical to what
  ``fr_weights_create(NULL, seed)`` produces (tests/test_oracle_weights.py pins that), so the oracle
  and the ``bench.py --impl reference`` arm get the shared weights WITHOUT loading libfr_b200.so.
* ``trained_like(model, seed)`` -- weights with the statistics trained checkpoints have and
  random-init does not: heavy-tailed (log-normal) BatchNorm scales, PReLU slopes over 0.01..0.9,
  a residual stream that grows to |activation| ~ 1e2, per-output-channel gain spread.  Layer
  gains are calibrated on a few probe inputs (LSUV-style) so the network stays finite.

Tensor names / shapes: oracle/nets.py ``det_tensor_specs`` / ``rec_tensor_specs``.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np

from . import nets

MODEL_DET, MODEL_REC = 0, 1
_M64 = (1 << 64) - 1


def _splitmix64_scalar(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


class _Filler:
    """csrc/weights.cpp ``Filler``: uniform in [-a, a) + offset, one splitmix64 draw per element."""

    def __init__(self, specs, seed: int):
        self.specs = specs
        self.index = {n: i for i, (n, _) in enumerate(specs)}
        self.seed = seed & _M64
        self.out: Dict[str, np.ndarray] = {}

    def sym(self, name: str, a: float, offset: float = 0.0):
        ti = self.index[name]
        shape = self.specs[ti][1]
        n = int(np.prod(shape))
        base = _splitmix64_scalar(self.seed ^ ((0xA24BAED4963EE407 * (ti + 1)) & _M64))
        with np.errstate(over="ignore"):
            h = _splitmix64(np.uint64(base) + np.arange(n, dtype=np.uint64))
        u = (h >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
        v = (np.float32(2.0) * u - np.float32(1.0)) * np.float32(a) + np.float32(offset)
        self.out[name] = v.astype(np.float32).reshape(shape)

    def var(self, name: str, v: float):
        self.sym(name, float(np.float32(math.sqrt(3.0 * v))))


def _seeded_det(seed: int) -> Dict[str, np.ndarray]:
    f = _Filler(nets.det_tensor_specs(), seed)

    def relu_conv(n, fan_in):
        f.var(n + ".w", 2.0 / fan_in)
        f.sym(n + ".b", 0.05)

    def lin_conv(n, fan_in):
        f.var(n + ".w", 1.0 / fan_in)
        f.sym(n + ".b", 0.05)

    def dwsep(n, ci):
        relu_conv(n + ".dw", 9)
        relu_conv(n + ".pw", ci)

    f.var("stem.w", 3.0 * 2.0 / 27.0)
    f.sym("stem.b", 0.05)
    dwsep("b0", 16)
    cin = 16
    for si, (nb, co) in enumerate(nets.DET_STAGES):
        for bi in range(nb):
            dwsep(f"s{si}.{bi}", cin)
            cin = co
    for i, c in enumerate((72, 152, 288)):
        lin_conv(f"lat{i}", c)
    for i in range(3):
        lin_conv(f"fpn{i}", 16 * 9)
    for i in range(2):
        lin_conv(f"down{i}", 16 * 9)
    for i in range(2):
        lin_conv(f"pafpn{i}", 16 * 9)
    for i in range(3):
        h = f"h{i}"
        dwsep(h + ".t0", 16)
        dwsep(h + ".t1", 64)
        f.var(h + ".cls.w", 9.0 / (64 * 9))
        f.sym(h + ".cls.b", 0.05, -4.595)
        f.var(h + ".reg.w", 0.25 / (64 * 9))
        f.sym(h + ".reg.b", 0.25, 1.5)
        f.var(h + ".kps.w", 1.0 / (64 * 9))
        f.sym(h + ".kps.b", 0.25, 0.0)
    return f.out


def _seeded_rec(seed: int) -> Dict[str, np.ndarray]:
    f = _Filler(nets.rec_tensor_specs(), seed)
    f.var("stem.w", 1.0 / (27.0 / 3.0))
    f.sym("stem.b", 0.05)
    f.sym("stem.prelu", 0.1, 0.25)
    v = 0.55
    cin = 64
    for li, (nb, planes) in enumerate(nets.REC_LAYERS):
        for bi in range(nb):
            p = f"l{li}.{bi}"
            f.sym(p + ".bn1.scale", float(np.float32(0.2 / math.sqrt(v))), float(np.float32(1.0 / math.sqrt(v))))
            f.sym(p + ".bn1.shift", 0.1)
            f.var(p + ".conv1.w", 1.0 / (9.0 * cin))
            f.sym(p + ".conv1.b", 0.05)
            f.sym(p + ".prelu", 0.1, 0.25)
            f.var(p + ".conv2.w", 0.5 / (9.0 * planes * 0.55))
            f.sym(p + ".conv2.b", 0.05)
            if bi == 0:
                f.var(p + ".ds.w", 1.0 / cin)
                f.sym(p + ".ds.b", 0.05)
            v += 0.5
            cin = planes
    f.sym("bn2.scale", float(np.float32(0.2 / math.sqrt(v))), float(np.float32(1.0 / math.sqrt(v))))
    f.sym("bn2.shift", 0.1)
    f.var("fc.w", 1.0 / 25088.0)
    f.sym("fc.b", 0.05)
    f.sym("feat.scale", 0.2, 1.0)
    f.sym("feat.shift", 0.1)
    return f.out


def seeded(model: int, seed: int) -> Dict[str, np.ndarray]:
    """Bit-identical to csrc/weights.cpp ``fr_weights_random_init(model, seed)``."""
    out = _seeded_det(seed) if model == MODEL_DET else _seeded_rec(seed)
    specs = nets.det_tensor_specs() if model == MODEL_DET else nets.rec_tensor_specs()
    return {n: out[n] for n, _ in specs}


# ------------------------------------------------------------------ trained-like --

def _lognormal(rng, n, sigma, clip=6.0):
    """Heavy-tailed positive gains with median 1."""
    return np.exp(np.clip(rng.normal(0.0, sigma, n), -clip * sigma, clip * sigma)).astype(np.float32)


def trained_like_rec(seed: int = 0, probe: int = 2, stream_growth: float = 1.13) -> Dict[str, np.ndarray]:
    """IResNet-50 weights with trained-checkpoint statistics.

    * PReLU slopes uniform in [0.01, 0.9]; 5 % of the channels at the ends of the range.
    * bn1 (pre-conv BatchNorm) scale = 1/std(stream) x log-normal(sigma 0.6), 3 % negative; shift N(0, 0.3).
    * conv rows carry a log-normal(sigma 0.5) per-output-channel gain (a folded BN gamma).
    * the residual branch of every block is calibrated (on `probe` random crops) so that the stream's
      std grows by `stream_growth` per block: |activation| passes 1e2 in layer 2 / 3.
    """
    import torch
    import torch.nn.functional as F

    rng = np.random.default_rng(seed)
    w: Dict[str, np.ndarray] = {}
    x = torch.from_numpy(((rng.integers(0, 256, (probe, 3, 112, 112)).astype(np.float32)) - 127.5) / 128)

    def conv_w(co, ci, k, gain_sigma=0.5):
        base = rng.standard_t(5, (co, ci, k, k)).astype(np.float32) * np.float32(math.sqrt(1.0 / (ci * k * k)) * 0.77)
        return base * _lognormal(rng, co, gain_sigma)[:, None, None, None]

    def slopes(c):
        s = rng.uniform(0.01, 0.9, c).astype(np.float32)
        m = rng.uniform(size=c)
        s[m < 0.025] = 0.01
        s[m > 0.975] = 0.9
        return s

    def t(a):
        return torch.from_numpy(a)

    with torch.no_grad():
        w["stem.w"] = conv_w(64, 3, 3) * np.float32(1.7)
        w["stem.b"] = rng.normal(0, 0.3, 64).astype(np.float32)
        w["stem.prelu"] = slopes(64)
        y = F.prelu(F.conv2d(x, t(w["stem.w"]), t(w["stem.b"]), padding=1), t(w["stem.prelu"]))
        cin = 64
        for li, (nb, planes) in enumerate(nets.REC_LAYERS):
            for bi in range(nb):
                p = f"l{li}.{bi}"
                stride = 2 if bi == 0 else 1
                sd = y.std(dim=(0, 2, 3)).numpy() + 1e-3
                mu = y.mean(dim=(0, 2, 3)).numpy()
                sc = (_lognormal(rng, cin, 0.6) / sd).astype(np.float32)
                sc[rng.uniform(size=cin) < 0.03] *= -1
                w[p + ".bn1.scale"] = sc
                w[p + ".bn1.shift"] = (rng.normal(0, 0.3, cin) - mu * sc).astype(np.float32)
                z = y * t(sc)[None, :, None, None] + t(w[p + ".bn1.shift"])[None, :, None, None]
                w[p + ".conv1.w"] = conv_w(planes, cin, 3)
                w[p + ".conv1.b"] = rng.normal(0, 0.4, planes).astype(np.float32)
                w[p + ".prelu"] = slopes(planes)
                z = F.prelu(F.conv2d(z, t(w[p + ".conv1.w"]), t(w[p + ".conv1.b"]), padding=1), t(w[p + ".prelu"]))
                w2 = conv_w(planes, planes, 3)
                b2 = rng.normal(0, 0.2, planes).astype(np.float32)
                z2 = F.conv2d(z, t(w2), None, stride=stride, padding=1)
                if bi == 0:
                    w[p + ".ds.w"] = conv_w(planes, cin, 1)
                    w[p + ".ds.b"] = rng.normal(0, 0.2, planes).astype(np.float32)
                    scut = F.conv2d(y, t(w[p + ".ds.w"]), t(w[p + ".ds.b"]), stride=stride)
                else:
                    scut = y
                target = float(scut.std()) * math.sqrt(stream_growth ** 2 - 1.0)
                g = np.float32(target / (float(z2.std()) + 1e-6))
                w[p + ".conv2.w"] = w2 * g
                w[p + ".conv2.b"] = b2 * np.float32(max(1.0, float(scut.std())) * 0.3)
                y = z2 * float(g) + t(w[p + ".conv2.b"])[None, :, None, None] + scut
                cin = planes
        sd = y.std(dim=(0, 2, 3)).numpy() + 1e-3
        mu = y.mean(dim=(0, 2, 3)).numpy()
        sc = (_lognormal(rng, 512, 0.4) / sd).astype(np.float32)
        w["bn2.scale"] = sc
        w["bn2.shift"] = (rng.normal(0, 0.2, 512) - mu * sc).astype(np.float32)
        w["fc.w"] = (rng.standard_t(5, (512, 25088)).astype(np.float32) * np.float32(0.77 / math.sqrt(25088.0)))
        w["fc.b"] = rng.normal(0, 0.05, 512).astype(np.float32)
        w["feat.scale"] = _lognormal(rng, 512, 0.3)
        w["feat.shift"] = rng.normal(0, 0.3, 512).astype(np.float32)
    return {n: np.ascontiguousarray(w[n], np.float32).reshape(s) for n, s in nets.rec_tensor_specs()}


def trained_like_det(seed: int = 0, probe: int = 1, pos_frac: float = 2.5e-3) -> Dict[str, np.ndarray]:
    """SCRFD weights with per-channel gain spread (log-normal sigma 0.5, a folded BN gamma), biases
    of the size trained BN shifts have, every ReLU layer calibrated to unit second moment on a
    probe frame, and the score bias set so that a fraction `pos_frac` of the anchors of each stride
    exceeds 0.5 on noise frames (so decode / NMS see a realistic candidate count)."""
    import torch
    import torch.nn.functional as F

    rng = np.random.default_rng(seed)
    w: Dict[str, np.ndarray] = {}
    specs = dict(nets.det_tensor_specs())
    for name, shape in specs.items():
        if name.endswith(".w"):
            co, ci, k, _ = shape
            fan = ci * k * k
            base = rng.standard_t(5, shape).astype(np.float32) * np.float32(0.77 * math.sqrt(2.0 / fan))
            w[name] = base * _lognormal(rng, co, 0.5)[:, None, None, None]
        else:
            w[name] = rng.normal(0, 0.15, shape).astype(np.float32)
    x = torch.from_numpy(((rng.integers(0, 256, (probe, 3, 640, 640)).astype(np.float32)) - 127.5) / 128)
    # calibrate layer by layer with the oracle's own forward (taps in canonical order)
    order = ["stem", "b0"] + [f"s{si}.{bi}" for si, (nb, _) in enumerate(nets.DET_STAGES) for bi in range(nb)]
    for ti, name in enumerate(order):
        _, taps = nets.scrfd_forward(w, x, return_taps=True)
        rms = float(taps[ti].pow(2).mean().sqrt())
        key = name + (".w" if name == "stem" else ".pw.w")
        g = np.float32(1.0 / max(rms, 1e-6))
        w[key] = w[key] * g
        w[key[:-1] + "b"] = w[key[:-1] + "b"] * g
    # neck (linear convs) and head towers: unit RMS at every tower output
    _, taps = nets.scrfd_forward(w, x, return_taps=True)
    nb = len(order)
    for i in range(3):                                   # laterals (taps nb .. nb+2 hold the merged maps)
        g = np.float32(1.0 / max(float(taps[nb + i].pow(2).mean().sqrt()), 1e-6))
        w[f"lat{i}.w"] *= g
        w[f"lat{i}.b"] *= g
    for tower in ("t0", "t1"):
        _, taps = nets.scrfd_forward(w, x, return_taps=True)
        base = len(taps) - (6 if tower == "t0" else 3)
        for i in range(3):
            g = np.float32(1.0 / max(float(taps[base + i].pow(2).mean().sqrt()), 1e-6))
            w[f"h{i}.{tower}.pw.w"] *= g
            w[f"h{i}.{tower}.pw.b"] *= g
    _, taps = nets.scrfd_forward(w, x, return_taps=True)
    for i in range(3):
        h = f"h{i}"
        t1 = taps[len(taps) - 3 + i]
        # distances / landmark offsets in stride units, sized like a trained head's (a face spans a few
        # strides; its landmarks lie inside the box): |bbox| <~ 3, |kps| <~ 3
        w[h + ".reg.w"] *= np.float32(0.2)
        w[h + ".reg.b"] = (rng.normal(1.5, 0.3, 8)).astype(np.float32)
        w[h + ".kps.w"] *= np.float32(0.2)
        w[h + ".kps.b"] = rng.normal(0, 0.3, 20).astype(np.float32)
        # score logits: std 1.5, bias at the (1 - pos_frac) quantile so that fraction exceeds 0.5
        logit = F.conv2d(t1, torch.from_numpy(w[h + ".cls.w"]), None, padding=1)
        g = np.float32(1.5 / max(float(logit.std()), 1e-6))
        w[h + ".cls.w"] *= g
        q = float(np.quantile((logit * float(g)).numpy().reshape(-1), 1.0 - pos_frac))
        w[h + ".cls.b"] = np.full(2, -q, np.float32)
    return {n: np.ascontiguousarray(w[n], np.float32).reshape(s) for n, s in nets.det_tensor_specs()}


def trained_like(model: int, seed: int = 0) -> Dict[str, np.ndarray]:
    return trained_like_det(seed) if model == MODEL_DET else trained_like_rec(seed)
