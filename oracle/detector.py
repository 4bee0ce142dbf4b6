"""CPU restatement of FaceDetector (reference src/face_detector.{h,cpp}).

TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import cv2
import numpy as np

from . import nets

INPUT_W = 640  # src/face_detector.cpp:8
INPUT_H = 640  # src/face_detector.cpp:9
STRIDES = (8, 16, 32)
NUM_ANCHORS = 2


@dataclass
class FaceBox:
    """struct FaceBox, src/face_detector.h:8-12 (cv::Rect is x,y,width,height ints)."""
    x: int
    y: int
    w: int
    h: int
    score: float
    landmarks: np.ndarray = field(default_factory=lambda: np.zeros((5, 2), np.float32))
    anchor: int = -1  # oracle-only bookkeeping: row index in the [N,15] tensor


def letterbox_geometry(rows: int, cols: int, in_w: int = INPUT_W, in_h: int = INPUT_H):
    """src/face_detector.cpp:101-106: fp32 scale, truncating int sizes."""
    scale_w = np.float32(in_w) / np.float32(cols)
    scale_h = np.float32(in_h) / np.float32(rows)
    scale = np.float32(min(scale_w, scale_h))
    new_w = int(np.float32(cols) * scale)  # static_cast<int> truncation
    new_h = int(np.float32(rows) * scale)
    return scale, new_w, new_h


def preprocess(image: np.ndarray, in_w: int = INPUT_W, in_h: int = INPUT_H):
    """FaceDetector::preprocess, src/face_detector.cpp:92-137.

    image: HxWx3 uint8 BGR.  Returns (chw float32 [3,in_h,in_w], scale) or
    (None, 1.0) on the reference's failure branches (:94-98, :109-113)."""
    if image is None or image.size == 0 or image.shape[0] <= 0 or image.shape[1] <= 0:
        return None, np.float32(1.0)
    rows, cols = image.shape[:2]
    scale, new_w, new_h = letterbox_geometry(rows, cols, in_w, in_h)
    if new_w <= 0 or new_h <= 0:
        return None, np.float32(1.0)
    resized = cv2.resize(image, (new_w, new_h))                 # :117
    padded = np.zeros((in_h, in_w, 3), np.uint8)                # :120
    padded[:new_h, :new_w] = resized                            # :121 (top-left paste)
    rgb = cv2.cvtColor(padded, cv2.COLOR_BGR2RGB)               # :125
    chw = (rgb.astype(np.float32) - np.float32(127.5)) / np.float32(128.0)  # :133
    return np.ascontiguousarray(chw.transpose(2, 0, 1)), scale


def scrfd_decode(heads: Sequence[np.ndarray], in_w: int = INPUT_W, in_h: int = INPUT_H) -> np.ndarray:
    """Three-stride anchor decode (absent from the reference; required by
    north_star; InsightFace scrfd.py distance2bbox/distance2kps semantics).

    heads: 9 arrays [N_s,1]x3, [N_s,4]x3, [N_s,10]x3 for strides 8,16,32 of
    ONE frame.  Produces the [N,15] tensor FaceDetector::postprocess consumes
    (src/face_detector.cpp:238-239,251,255-258,271-272):
    x1,y1,x2,y2,score,kx0,ky0..kx4,ky4 in network-input (640) pixels.  All
    arithmetic is fp32 with separate multiply and add/sub roundings."""
    rows = []
    for si, s in enumerate(STRIDES):
        hs, ws = in_h // s, in_w // s
        sc = np.asarray(heads[si], np.float32).reshape(-1)
        bb = np.asarray(heads[3 + si], np.float32).reshape(-1, 4)
        kp = np.asarray(heads[6 + si], np.float32).reshape(-1, 10)
        n = hs * ws * NUM_ANCHORS
        assert sc.shape[0] == n and bb.shape[0] == n and kp.shape[0] == n
        idx = np.arange(n) // NUM_ANCHORS
        cx = ((idx % ws) * s).astype(np.float32)
        cy = ((idx // ws) * s).astype(np.float32)
        st = np.float32(s)
        out = np.empty((n, 15), np.float32)
        out[:, 0] = cx - bb[:, 0] * st
        out[:, 1] = cy - bb[:, 1] * st
        out[:, 2] = cx + bb[:, 2] * st
        out[:, 3] = cy + bb[:, 3] * st
        out[:, 4] = sc
        for j in range(5):
            out[:, 5 + 2 * j] = cx + kp[:, 2 * j] * st
            out[:, 6 + 2 * j] = cy + kp[:, 2 * j + 1] * st
        rows.append(out)
    return np.concatenate(rows, 0)


def _trunc_i32(v) -> int:
    """static_cast<int>(float): truncation toward zero.  Out-of-range / NaN is
    UB in C++; we saturate like the CUDA cvt.rzi.s32.f32 the product uses."""
    v = float(v)
    if v != v:
        return 0
    if v >= 2147483647.0:
        return 2147483647
    if v <= -2147483648.0:
        return -2147483648
    return int(v)


def iou(a: FaceBox, b: FaceBox) -> float:
    """FaceDetector::iou, src/face_detector.cpp:340-354: all-int arithmetic,
    float(inter) / float(int denominator); 0/0 -> NaN (never '>' threshold)."""
    x1 = max(a.x, b.x)
    y1 = max(a.y, b.y)
    x2 = min(a.x + a.w, b.x + b.w)
    y2 = min(a.y + a.h, b.y + b.h)
    w = max(0, x2 - x1)
    h = max(0, y2 - y1)
    inter = np.int32(w) * np.int32(h)
    area1 = np.int32(a.w) * np.int32(a.h)
    area2 = np.int32(b.w) * np.int32(b.h)
    den = np.int32(area1 + area2 - inter)
    with np.errstate(all="ignore"):
        return np.float32(inter) / np.float32(den)


def nms(boxes: List[FaceBox], threshold: float) -> List[FaceBox]:
    """FaceDetector::nms, src/face_detector.cpp:356-384.  std::sort is
    unstable; the canonical tie order is score desc, then
    candidate (anchor) index asc.  The inner j-loop is vectorised with numpy
    int32 / float32 arrays (same arithmetic as ``iou`` element-wise)."""
    order = sorted(range(len(boxes)), key=lambda i: (-float(boxes[i].score), boxes[i].anchor, i))
    bs = [boxes[i] for i in order]
    n = len(bs)
    if n == 0:
        return []
    thr = np.float32(threshold)
    X = np.array([b.x for b in bs], np.int32)
    Y = np.array([b.y for b in bs], np.int32)
    W = np.array([b.w for b in bs], np.int32)
    H = np.array([b.h for b in bs], np.int32)
    with np.errstate(all="ignore"):
        X2, Y2, A = X + W, Y + H, W * H
        suppressed = np.zeros(n, bool)
        for i in range(n):
            if suppressed[i]:
                continue
            j = slice(i + 1, n)
            w = np.maximum(np.int32(0), np.minimum(X2[i], X2[j]) - np.maximum(X[i], X[j]))
            h = np.maximum(np.int32(0), np.minimum(Y2[i], Y2[j]) - np.maximum(Y[i], Y[j]))
            inter = (w * h).astype(np.int32)
            den = (A[i] + A[j] - inter).astype(np.int32)
            v = inter.astype(np.float32) / den.astype(np.float32)
            suppressed[j] |= v > thr  # strict, :370; NaN (0/0) compares false
    return [b for b, s_ in zip(bs, suppressed) if not s_]


def postprocess(out15: np.ndarray, scale, score_thr: float = 0.5, nms_thr: float = 0.4) -> List[FaceBox]:
    """FaceDetector::postprocess, src/face_detector.cpp:224-338 ([N,>=15] branch)."""
    out15 = np.asarray(out15, np.float32)
    if out15.ndim == 3:  # [1,N,F]: batch 0 only (:244-250)
        out15 = out15[0]
    scale = np.float32(scale)
    thr = np.float32(score_thr)
    boxes: List[FaceBox] = []
    for i in np.nonzero(out15[:, 4] > thr)[0]:  # strict '>' (:253)
        o = out15[i]
        x1, y1, x2, y2 = (np.float32(o[k]) / scale for k in range(4))  # :255-258
        lm = (o[5:15] / scale).astype(np.float32).reshape(5, 2)        # :270-273
        boxes.append(FaceBox(_trunc_i32(x1), _trunc_i32(y1),
                             _trunc_i32(np.float32(x2 - x1)), _trunc_i32(np.float32(y2 - y1)),
                             float(o[4]), lm, int(i)))                 # :260-265
    return nms(boxes, nms_thr)                                         # :333


class FaceDetector:
    """Mirror of class FaceDetector (src/face_detector.h:14-43) on CPU."""

    def __init__(self, weights=None):
        self.weights = weights
        self.input_w, self.input_h = INPUT_W, INPUT_H

    def load_weights(self, weights) -> bool:
        self.weights = weights
        return True

    def run_network(self, chw: np.ndarray):
        import torch
        outs = nets.scrfd_forward(self.weights, torch.from_numpy(chw[None]))
        return [o[0].numpy() for o in outs]

    def detect(self, image: np.ndarray, score_thr: float = 0.5, nms_thr: float = 0.4) -> List[FaceBox]:
        """FaceDetector::detect, src/face_detector.cpp:139-222."""
        if self.weights is None or image is None or image.size == 0:
            return []
        chw, scale = preprocess(image, self.input_w, self.input_h)
        if chw is None:
            return []
        heads = self.run_network(chw)
        return postprocess(scrfd_decode(heads, self.input_w, self.input_h), scale, score_thr, nms_thr)
