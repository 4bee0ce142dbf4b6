"""torch-CPU fp32 restatement of the two ``Ort::Session::Run`` calls.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference executes two fixed ONNX graphs through ONNX Runtime's CPU EP
(src/face_detector.cpp:179-183, src/face_recognizer.cpp:196-200,279-283).
Neither graph nor ONNX Runtime is present in /root/reference or this image,
so the architectures are restated from the published InsightFace definitions
(SCRFD ``scrfd_500m_bnkps``; arcface_torch ``iresnet50``) in the *exported*
parameterisation: every conv->BN pair folded into conv+bias, the pre-conv
BatchNorm of each IBasicBlock and the tail BatchNorms kept as per-channel
affine (scale, shift).  I/O contracts follow models/README.md:12,18-19.

Weights are a ``dict[str, np.ndarray]`` in canonical layouts (conv OIHW,
fc [out,in]); the tensor list (names, shapes, order) is ``det_tensor_specs()``
/ ``rec_tensor_specs()`` and must equal what the C-ABI reports through
``fr_weights_tensor_info`` (tests/test_capi_host.py checks that).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# ------------------------------------------------------------------ SCRFD --

DET_STAGES = [(2, 40), (3, 72), (2, 152), (6, 288)]  # (blocks, out channels); first block stride 2
DET_NECK_C = 16
DET_HEAD_C = 64
DET_STRIDES = (8, 16, 32)
DET_NUM_ANCHORS = 2


def det_tensor_specs() -> List[Tuple[str, Tuple[int, ...]]]:
    specs: List[Tuple[str, Tuple[int, ...]]] = []

    def conv(name, co, ci, k):
        specs.append((name + ".w", (co, ci, k, k)))
        specs.append((name + ".b", (co,)))

    def dwsep(name, ci, co):
        conv(name + ".dw", ci, 1, 3)
        conv(name + ".pw", co, ci, 1)

    conv("stem", 16, 3, 3)
    dwsep("b0", 16, 16)
    cin = 16
    for si, (nb, co) in enumerate(DET_STAGES):
        for bi in range(nb):
            dwsep(f"s{si}.{bi}", cin, co)
            cin = co
    feats = [DET_STAGES[1][1], DET_STAGES[2][1], DET_STAGES[3][1]]
    for i, c in enumerate(feats):
        conv(f"lat{i}", DET_NECK_C, c, 1)
    for i in range(3):
        conv(f"fpn{i}", DET_NECK_C, DET_NECK_C, 3)
    for i in range(2):
        conv(f"down{i}", DET_NECK_C, DET_NECK_C, 3)
    for i in range(2):
        conv(f"pafpn{i}", DET_NECK_C, DET_NECK_C, 3)
    for i in range(3):
        dwsep(f"h{i}.t0", DET_NECK_C, DET_HEAD_C)
        dwsep(f"h{i}.t1", DET_HEAD_C, DET_HEAD_C)
        conv(f"h{i}.cls", DET_NUM_ANCHORS * 1, DET_HEAD_C, 3)
        conv(f"h{i}.reg", DET_NUM_ANCHORS * 4, DET_HEAD_C, 3)
        conv(f"h{i}.kps", DET_NUM_ANCHORS * 10, DET_HEAD_C, 3)
    return specs


_DTYPE = np.float32   # float64 inside ``reference_precision()``: the yardstick two fp32 engines are judged by


class reference_precision:
    """Context manager: run the restated graphs in float64 (inputs must be float64 tensors).  Used by
    the tests to separate an engine's own rounding error from a real defect: an fp32 engine's distance
    to this result is the noise floor every other fp32 engine is allowed."""

    def __enter__(self):
        global _DTYPE
        self._old, _DTYPE = _DTYPE, np.float64
        return self

    def __exit__(self, *exc):
        global _DTYPE
        _DTYPE = self._old
        return False


def _t(w: Dict[str, np.ndarray], name: str) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(w[name], dtype=_DTYPE))


def _conv(w, name, x, stride=1, relu=False, groups=1):
    k = w[name + ".w"].shape[-1]
    y = F.conv2d(x, _t(w, name + ".w"), _t(w, name + ".b"), stride=stride, padding=k // 2, groups=groups)
    return F.relu(y) if relu else y


def _dwsep(w, name, x, stride):
    c = x.shape[1]
    y = _conv(w, name + ".dw", x, stride=stride, relu=True, groups=c)
    return _conv(w, name + ".pw", y, relu=True)


@torch.no_grad()
def scrfd_forward(w: Dict[str, np.ndarray], x: torch.Tensor, return_taps: bool = False):
    """x: [B,3,H,W] fp32 RGB in [-1,1].  Returns the 9 head tensors in the
    order/layout of the buffalo export, batched: scores [B,N_s,1], bbox
    [B,N_s,4], kps [B,N_s,10] for strides 8,16,32 (anchor index =
    (gy*W_s+gx)*2+a; distances/offsets in stride units; sigmoid on scores).
    With return_taps also a list of the intermediate activations (NCHW) in the
    order of the C-ABI test hook fr_scrfd_tap."""
    taps = []
    y = _conv(w, "stem", x, stride=2, relu=True)
    taps.append(y)
    y = _dwsep(w, "b0", y, 1)
    taps.append(y)
    feats = []
    for si, (nb, _) in enumerate(DET_STAGES):
        for bi in range(nb):
            y = _dwsep(w, f"s{si}.{bi}", y, 2 if bi == 0 else 1)
            taps.append(y)
        if si >= 1:
            feats.append(y)
    lat = [_conv(w, f"lat{i}", f) for i, f in enumerate(feats)]
    for i in (2, 1):
        lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], scale_factor=2, mode="nearest")
    taps += lat
    inter = [_conv(w, f"fpn{i}", lat[i]) for i in range(3)]
    for i in range(2):
        inter[i + 1] = inter[i + 1] + _conv(w, f"down{i}", inter[i], stride=2)
    taps += inter
    outs = [inter[0]] + [_conv(w, f"pafpn{i - 1}", inter[i]) for i in (1, 2)]
    taps += outs[1:]
    scores, bboxes, kpss = [], [], []
    t0s, t1s = [], []
    for i, f in enumerate(outs):
        t = _dwsep(w, f"h{i}.t0", f, 1)
        t0s.append(t)
        t = _dwsep(w, f"h{i}.t1", t, 1)
        t1s.append(t)
        B = t.shape[0]
        cls = torch.sigmoid(_conv(w, f"h{i}.cls", t)).permute(0, 2, 3, 1).reshape(B, -1, 1)
        reg = _conv(w, f"h{i}.reg", t).permute(0, 2, 3, 1).reshape(B, -1, 4)
        kps = _conv(w, f"h{i}.kps", t).permute(0, 2, 3, 1).reshape(B, -1, 10)
        scores.append(cls)
        bboxes.append(reg)
        kpss.append(kps)
    taps += t0s + t1s
    if return_taps:
        return scores + bboxes + kpss, taps
    return scores + bboxes + kpss


# --------------------------------------------------------------- IResNet-50 --

REC_LAYERS = [(3, 64), (4, 128), (14, 256), (3, 512)]
REC_FEAT = 512


def rec_tensor_specs() -> List[Tuple[str, Tuple[int, ...]]]:
    specs: List[Tuple[str, Tuple[int, ...]]] = [
        ("stem.w", (64, 3, 3, 3)), ("stem.b", (64,)), ("stem.prelu", (64,))]
    cin = 64
    for li, (nb, planes) in enumerate(REC_LAYERS):
        for bi in range(nb):
            p = f"l{li}.{bi}"
            specs += [(p + ".bn1.scale", (cin,)), (p + ".bn1.shift", (cin,)),
                      (p + ".conv1.w", (planes, cin, 3, 3)), (p + ".conv1.b", (planes,)),
                      (p + ".prelu", (planes,)),
                      (p + ".conv2.w", (planes, planes, 3, 3)), (p + ".conv2.b", (planes,))]
            if bi == 0:
                specs += [(p + ".ds.w", (planes, cin, 1, 1)), (p + ".ds.b", (planes,))]
            cin = planes
    specs += [("bn2.scale", (512,)), ("bn2.shift", (512,)),
              ("fc.w", (REC_FEAT, 512 * 7 * 7)), ("fc.b", (REC_FEAT,)),
              ("feat.scale", (REC_FEAT,)), ("feat.shift", (REC_FEAT,))]
    return specs


@torch.no_grad()
def iresnet50_forward(w: Dict[str, np.ndarray], x: torch.Tensor, return_taps: bool = False):
    """x: [B,3,112,112] fp32 RGB in [-1,1] -> [B,512] (NOT yet L2-normalised;
    the reference normalises outside the graph, src/face_recognizer.cpp:297)."""
    taps = {}
    y = F.conv2d(x, _t(w, "stem.w"), _t(w, "stem.b"), padding=1)
    y = F.prelu(y, _t(w, "stem.prelu"))
    taps["stem"] = y
    for li, (nb, planes) in enumerate(REC_LAYERS):
        for bi in range(nb):
            p = f"l{li}.{bi}"
            stride = 2 if bi == 0 else 1
            z = y * _t(w, p + ".bn1.scale")[None, :, None, None] + _t(w, p + ".bn1.shift")[None, :, None, None]
            z = F.conv2d(z, _t(w, p + ".conv1.w"), _t(w, p + ".conv1.b"), padding=1)
            z = F.prelu(z, _t(w, p + ".prelu"))
            if return_taps:
                taps[p + ".h"] = z
            z = F.conv2d(z, _t(w, p + ".conv2.w"), _t(w, p + ".conv2.b"), stride=stride, padding=1)
            if bi == 0:
                sc = F.conv2d(y, _t(w, p + ".ds.w"), _t(w, p + ".ds.b"), stride=stride)
            else:
                sc = y
            y = z + sc
            if return_taps:
                taps[p] = y
    y = y * _t(w, "bn2.scale")[None, :, None, None] + _t(w, "bn2.shift")[None, :, None, None]
    y = y.flatten(1)
    y = F.linear(y, _t(w, "fc.w"), _t(w, "fc.b"))
    y = y * _t(w, "feat.scale")[None] + _t(w, "feat.shift")[None]
    if return_taps:
        return y, taps
    return y
