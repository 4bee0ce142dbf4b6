"""CPU restatement of FaceRecognizer (reference src/face_recognizer.{h,cpp}).

TEST INFRASTRUCTURE (see oracle/__init__.py).
"""
from __future__ import annotations

from typing import Optional

import cv2
import numpy as np

from . import nets
from .detector import FaceBox

INPUT_W = 112   # src/face_recognizer.cpp:8
INPUT_H = 112   # src/face_recognizer.cpp:9
FEATURE_DIM = 512  # src/face_recognizer.cpp:10

# src/face_recognizer.cpp:101-107
TEMPLATE = np.array([[38.2946, 51.6963], [73.5318, 51.5014], [56.0252, 71.7366],
                     [41.5493, 92.3655], [70.7299, 92.2041]], np.float32)

SAME_PERSON_THRESHOLD = np.float32(0.6)  # src/main.cpp:118


def estimate_alignment(landmarks: np.ndarray) -> Optional[np.ndarray]:
    """cv::estimateAffinePartial2D(landmarks -> template), all defaults
    (src/face_recognizer.cpp:110-113).  Returns 2x3 float64 or None."""
    M, _ = cv2.estimateAffinePartial2D(np.asarray(landmarks, np.float32).reshape(5, 2), TEMPLATE)
    return M


def align_face(image: np.ndarray, face: FaceBox) -> Optional[np.ndarray]:
    """FaceRecognizer::alignFace, src/face_recognizer.cpp:93-133."""
    if image is None or image.size == 0:
        return None
    M = estimate_alignment(face.landmarks)
    if M is None:
        # :116-127  crop (box & image) and plain resize
        rows, cols = image.shape[:2]
        x0, y0 = max(face.x, 0), max(face.y, 0)
        x1, y1 = min(face.x + face.w, cols), min(face.y + face.h, rows)
        if x1 - x0 > 0 and y1 - y0 > 0:
            return cv2.resize(np.ascontiguousarray(image[y0:y1, x0:x1]), (INPUT_W, INPUT_H))
        return None
    return cv2.warpAffine(image, M, (INPUT_W, INPUT_H))  # :130


def preprocess(aligned: np.ndarray) -> np.ndarray:
    """FaceRecognizer::preprocess, src/face_recognizer.cpp:135-150."""
    rgb = cv2.cvtColor(aligned, cv2.COLOR_BGR2RGB)
    chw = (rgb.astype(np.float32) - np.float32(127.5)) / np.float32(128.0)
    return np.ascontiguousarray(chw.transpose(2, 0, 1))


def normalize(feature: np.ndarray) -> np.ndarray:
    """FaceRecognizer::normalize, src/face_recognizer.cpp:306-318:
    sequential fp32 sum of squares, sqrt, divide iff norm > 0."""
    f = np.asarray(feature, np.float32).copy()
    norm = np.float32(0.0)
    for v in f:
        norm = np.float32(norm + np.float32(v * v))
    norm = np.float32(np.sqrt(norm))
    if norm > 0:
        f = (f / norm).astype(np.float32)
    return f


def normalize_rows(feats: np.ndarray) -> np.ndarray:
    """Vectorised variant for batches (pairwise summation; low-bit differences
    only, well inside the cosine >= 0.999 budget)."""
    f = np.asarray(feats, np.float32)
    n = np.sqrt((f * f).sum(axis=1, dtype=np.float32)).astype(np.float32)
    out = f.copy()
    nz = n > 0
    out[nz] = f[nz] / n[nz, None]
    return out


def compare_faces(f1: np.ndarray, f2: np.ndarray) -> np.float32:
    """FaceRecognizer::compareFaces, src/face_recognizer.cpp:320-334."""
    f1 = np.asarray(f1, np.float32).reshape(-1)
    f2 = np.asarray(f2, np.float32).reshape(-1)
    if f1.size != f2.size or f1.size == 0:
        return np.float32(0.0)
    dot = np.float32(0.0)
    for a, b in zip(f1, f2):
        dot = np.float32(dot + np.float32(a * b))
    return np.float32(np.float32(dot + np.float32(1.0)) / np.float32(2.0))


def same_person(similarity) -> bool:
    """src/main.cpp:118-123 decision rule (strict '>')."""
    return bool(np.float32(similarity) > SAME_PERSON_THRESHOLD)


class FaceRecognizer:
    """Mirror of class FaceRecognizer (src/face_recognizer.h:9-38) on CPU."""

    def __init__(self, weights=None):
        self.weights = weights

    def load_weights(self, weights) -> bool:
        self.weights = weights
        return True

    def embed_chw(self, chw_batch: np.ndarray) -> np.ndarray:
        import torch
        out = nets.iresnet50_forward(self.weights, torch.from_numpy(np.ascontiguousarray(chw_batch, np.float32)))
        return out.numpy()

    def extract_feature(self, image: np.ndarray, face: FaceBox) -> np.ndarray:
        """FaceRecognizer::extractFeature, src/face_recognizer.cpp:236-304."""
        if self.weights is None or image is None or image.size == 0:
            return np.zeros(0, np.float32)
        aligned = align_face(image, face)
        if aligned is None:
            return np.zeros(0, np.float32)
        return normalize(self.embed_chw(preprocess(aligned)[None])[0])

    def extract_feature_simple(self, image: np.ndarray) -> np.ndarray:
        """FaceRecognizer::extractFeatureSimple, src/face_recognizer.cpp:152-234."""
        if self.weights is None or image is None or image.size == 0:
            return np.zeros(0, np.float32)
        resized = cv2.resize(image, (INPUT_W, INPUT_H))
        return normalize(self.embed_chw(preprocess(resized)[None])[0])

    compare_faces = staticmethod(compare_faces)
