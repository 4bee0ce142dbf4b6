"""Python mirror of the reference's public classes, on top of the C ABI.

Same names, argument meaning and error behaviour as the reference:

* ``FaceDetector.loadModel(path) -> bool`` / ``detect(image, scoreThreshold=0.5,
  nmsThreshold=0.4) -> list[FaceBox]``   (src/face_detector.h:16-20)
* ``FaceRecognizer.loadModel(path) -> bool`` / ``extractFeature(image, face)`` /
  ``extractFeatureSimple(image)`` / ``compareFaces(f1, f2)``   (src/face_recognizer.h:11-17)

Failures print the reference's message on stderr and return ``[]`` / empty array / ``0.0``
(src/face_detector.cpp:142-167,217-219; src/face_recognizer.cpp:321-323).  When the model
file is absent the weights are seeded random-init of the same architecture and loadModel
returns True with a notice (BASELINE.json north_star).
"""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import capi


@dataclass
class FaceBox:
    """struct FaceBox (src/face_detector.h:8-12): box = (x, y, width, height) ints."""
    box: tuple
    score: float
    landmarks: np.ndarray = field(default_factory=lambda: np.zeros((5, 2), np.float32))

    def to_record(self) -> np.ndarray:
        r = np.zeros(1, capi.FACE_DTYPE)
        r["x"], r["y"], r["w"], r["h"] = self.box
        r["score"] = self.score
        r["lm"] = np.asarray(self.landmarks, np.float32).reshape(10)
        return r

    @staticmethod
    def from_record(r) -> "FaceBox":
        return FaceBox((int(r["x"]), int(r["y"]), int(r["w"]), int(r["h"])), float(r["score"]),
                       np.array(r["lm"], np.float32).reshape(5, 2))


class _Shared:
    """One device context shared by a detector and a recognizer created on the same device."""
    ctx: Optional[capi.Context] = None
    det_w: Optional[capi.Weights] = None
    rec_w: Optional[capi.Weights] = None
    device: int = 0
    seed: int = 0
    # random-init weights when the model file is absent are an explicit opt-in (as in host/face_api.cpp)
    allow_random_init: bool = os.environ.get("FR_ALLOW_RANDOM_INIT", "0") not in ("", "0")

    @classmethod
    def rebuild(cls):
        if cls.ctx is not None:
            cls.ctx.close()
        cls.ctx = capi.Context(cls.device, cls.det_w, cls.rec_w)


def _load(model: int, path: str, what: str) -> Optional[capi.Weights]:
    if path and os.path.exists(path):
        try:
            return capi.Weights(model, path)
        except capi.FrError as e:
            print(f"Error loading {what} model: {e}", file=sys.stderr)
            return None
    if not _Shared.allow_random_init:
        # the reference returns false here (src/face_detector.cpp:86-89); so do we
        print(f"Error loading {what} model: cannot open {path} "
              "(set FR_ALLOW_RANDOM_INIT=1 to run with seeded random-init weights)", file=sys.stderr)
        return None
    print(f"WARNING: {path} not found; FR_ALLOW_RANDOM_INIT=1 -> seeded random-init {what} weights of the "
          "same architecture (results are meaningless for real faces)", file=sys.stderr)
    return capi.Weights(model, None, _Shared.seed)


class FaceDetector:
    def __init__(self):
        self._loaded = False

    def loadModel(self, modelPath: str) -> bool:
        w = _load(capi.FR_MODEL_DET, modelPath, "face detector")
        if w is None:
            return False
        _Shared.det_w = w
        try:
            _Shared.rebuild()
        except capi.FrError as e:
            print(f"Error loading face detector model: {e}", file=sys.stderr)
            return False
        self._loaded = True
        return True

    def detect(self, image: np.ndarray, scoreThreshold: float = 0.5, nmsThreshold: float = 0.4) -> List[FaceBox]:
        if not self._loaded:
            print("Model not loaded!", file=sys.stderr)
            return []
        if image is None or image.size == 0:
            print("Input image is empty!", file=sys.stderr)
            return []
        try:
            recs = _Shared.ctx.detect(image, scoreThreshold, nmsThreshold)
        except capi.FrError as e:
            print(f"Error during inference: {e}", file=sys.stderr)
            return []
        return [FaceBox.from_record(r) for r in recs]


class FaceRecognizer:
    def __init__(self):
        self._loaded = False

    def loadModel(self, modelPath: str) -> bool:
        w = _load(capi.FR_MODEL_REC, modelPath, "face recognizer")
        if w is None:
            return False
        _Shared.rec_w = w
        try:
            _Shared.rebuild()
        except capi.FrError as e:
            print(f"Error loading face recognizer model: {e}", file=sys.stderr)
            return False
        self._loaded = True
        return True

    def extractFeature(self, image: np.ndarray, face: FaceBox) -> np.ndarray:
        if not self._loaded:
            print("Model not loaded!", file=sys.stderr)
            return np.zeros(0, np.float32)
        if image is None or image.size == 0:
            print("Input image is empty!", file=sys.stderr)
            return np.zeros(0, np.float32)
        try:
            emb, valid = _Shared.ctx.embed_faces([image], face.to_record(), [0])
        except capi.FrError as e:
            print(f"Error during feature extraction: {e}", file=sys.stderr)
            return np.zeros(0, np.float32)
        if not valid[0]:
            print("Face alignment failed!", file=sys.stderr)
            return np.zeros(0, np.float32)
        return emb[0]

    def extractFeatureSimple(self, image: np.ndarray) -> np.ndarray:
        if not self._loaded:
            print("Model not loaded!", file=sys.stderr)
            return np.zeros(0, np.float32)
        if image is None or image.size == 0:
            print("Input image is empty!", file=sys.stderr)
            return np.zeros(0, np.float32)
        try:
            return _Shared.ctx.embed_simple(image)
        except capi.FrError as e:
            print(f"Error during feature extraction: {e}", file=sys.stderr)
            return np.zeros(0, np.float32)

    def compareFaces(self, feature1, feature2) -> float:
        return capi.compare(np.asarray(feature1, np.float32), np.asarray(feature2, np.float32))
