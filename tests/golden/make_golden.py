"""Generates the committed fixtures under tests/golden/.

Run from the repo root:  python tests/golden/make_golden.py

* ref_arith.json : HAND-DERIVED known-answer cases for the reference-owned arithmetic
  (src/face_detector.cpp:253-265 truncation, :340-354 integer IoU, :356-384 greedy NMS,
  src/face_recognizer.cpp:306-334 normalize / compareFaces).  The expected values are
  written out literally below (derived by hand from the C++ source), NOT produced by the
  oracle -- they pin the oracle.
* decode_nms_case.npz : seeded synthetic SCRFD head tensors for 2 frames plus the oracle's
  kept detections (regression fixture for the CUDA decode+NMS path).
* align_case.npz : a seeded 320x240 image, 6 landmark sets and cv2's aligned crops.
The reference itself ships no fixtures and cannot run here.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import detector as odet  # noqa: E402
from oracle import recognizer as orec  # noqa: E402


def ref_arith():
    cases = {
        # cv::Rect(x,y,w,h) pairs -> float(inter)/(a1+a2-inter)   (face_detector.cpp:340-354)
        "iou": [
            {"a": [0, 0, 10, 10], "b": [5, 5, 10, 10], "num": 25, "den": 175},
            {"a": [0, 0, 10, 10], "b": [0, 0, 10, 10], "num": 100, "den": 100},
            {"a": [0, 0, 10, 10], "b": [10, 0, 10, 10], "num": 0, "den": 200},
            {"a": [0, 0, 10, 10], "b": [20, 20, 5, 5], "num": 0, "den": 125},
            {"a": [2, 3, 7, 4], "b": [4, 1, 10, 5], "num": 15, "den": 63},
            {"a": [0, 0, 0, 0], "b": [0, 0, 0, 0], "num": 0, "den": 0},          # 0/0 -> NaN -> never suppresses
            {"a": [0, 0, -4, 5], "b": [-3, 0, 10, 5], "num": 0, "den": 30},        # negative width: w = max(0, -4 - 0)
            {"a": [-5, -5, 10, 10], "b": [0, 0, 10, 10], "num": 25, "den": 175},
        ],
        # rows [x1,y1,x2,y2] / scale -> cv::Rect(int(x1), int(y1), int(x2-x1), int(y2-y1))   (:255-265)
        "rect": [
            {"row": [10.9, 20.1, 50.5, 80.9], "scale": 1.0, "rect": [10, 20, 39, 60]},
            {"row": [-3.7, -0.2, 4.2, 9.9], "scale": 1.0, "rect": [-3, 0, 7, 10]},      # trunc toward zero
            {"row": [100.0, 50.0, 300.0, 250.0], "scale": 0.5, "rect": [200, 100, 400, 400]},
            {"row": [64.0, 32.0, 96.0, 160.0], "scale": 2.0, "rect": [32, 16, 16, 64]},
            {"row": [9.0, 9.0, 3.0, 3.0], "scale": 1.0, "rect": [9, 9, -6, -6]},        # inverted box keeps sign
        ],
        # greedy NMS, boxes as [x,y,w,h,score]; expected = indices kept, in output order   (:356-384)
        "nms": [
            {"boxes": [[0, 0, 10, 10, 0.9], [1, 1, 10, 10, 0.8], [50, 50, 10, 10, 0.7]], "thr": 0.4,
             "keep": [0, 2]},                                    # iou(0,1)=81/119=0.68 > 0.4
            {"boxes": [[0, 0, 10, 10, 0.6], [5, 0, 10, 10, 0.9], [10, 0, 10, 10, 0.8]], "thr": 0.4,
             "keep": [1, 2, 0]},                                 # iou=50/150=0.333 (not > 0.4): all kept, sorted
            {"boxes": [[0, 0, 10, 10, 0.9], [5, 0, 10, 10, 0.8], [5, 0, 10, 10, 0.7]], "thr": 0.3,
             "keep": [0]},                                       # both later boxes have iou 1/3 > 0.3 with box 0
            {"boxes": [[0, 0, 10, 10, 0.9], [0, 0, 10, 10, 0.8], [0, 0, 10, 10, 0.85]], "thr": 1.0,
             "keep": [0, 2, 1]},                                 # iou == 1.0 is NOT > 1.0 (strict)
            {"boxes": [[0, 0, 0, 0, 0.9], [0, 0, 0, 0, 0.8]], "thr": 0.0, "keep": [0, 1]},   # NaN never suppresses
            {"boxes": [[0, 0, 10, 10, 0.9], [2, 0, 10, 10, 0.8], [4, 0, 10, 10, 0.7]], "thr": 0.5,
             "keep": [0, 2]},                                    # 0 kills 1 (80/120); 0 vs 2 = 60/140 = .43 kept;
        ],                                                       # 1 is dead so it cannot kill 2
        # compareFaces (face_recognizer.cpp:320-334)
        "compare": [
            {"a": [1.0, 0.0, 0.0], "b": [1.0, 0.0, 0.0], "sim": 1.0},
            {"a": [1.0, 0.0, 0.0], "b": [-1.0, 0.0, 0.0], "sim": 0.0},
            {"a": [1.0, 0.0, 0.0], "b": [0.0, 1.0, 0.0], "sim": 0.5},
            {"a": [0.6, 0.8], "b": [0.8, 0.6], "sim": 0.98},     # dot = .96 -> (1.96)/2
            {"a": [1.0, 0.0], "b": [1.0, 0.0, 0.0], "sim": 0.0},  # size mismatch -> 0.0f
            {"a": [], "b": [], "sim": 0.0},                       # empty -> 0.0f
        ],
        # normalize (face_recognizer.cpp:306-318)
        "normalize": [
            {"v": [3.0, 4.0], "out": [0.6, 0.8]},
            {"v": [0.0, 0.0, 0.0], "out": [0.0, 0.0, 0.0]},       # norm == 0: untouched
            {"v": [2.0, 0.0, 0.0, 0.0], "out": [1.0, 0.0, 0.0, 0.0]},
        ],
        # main.cpp:118-123: same person iff similarity > 0.6f (strict)
        "decision": [{"sim": 0.6, "same": False}, {"sim": 0.60001, "same": True}, {"sim": 0.59, "same": False}],
    }
    with open(os.path.join(HERE, "ref_arith.json"), "w") as f:
        json.dump(cases, f, indent=1)


def decode_nms_case():
    rng = np.random.default_rng(2024)
    n = 2
    heads = []
    for k, c in enumerate((1, 4, 10)):
        for s, ns in enumerate((12800, 3200, 800)):
            if k == 0:
                z = rng.normal(-3.2, 1.2, (n, ns, c))
                a = 1.0 / (1.0 + np.exp(-z))
                # exact-tie scores (sigmoid saturation analogue) to exercise the canonical tie order
                a[:, ::997, :] = 0.75
            elif k == 1:
                a = rng.normal(1.8, 0.8, (n, ns, c))
            else:
                a = rng.normal(0.0, 1.0, (n, ns, c))
            heads.append(a.astype(np.float32))
    scales = np.array([1.0, 0.5], np.float32)
    out = {f"head{i}": h for i, h in enumerate(heads)}
    out["scales"] = scales
    for i in range(n):
        faces = odet.postprocess(odet.scrfd_decode([h[i] for h in heads]), scales[i], 0.5, 0.4)
        out[f"rect{i}"] = np.array([[f.x, f.y, f.w, f.h] for f in faces], np.int32).reshape(-1, 4)
        out[f"score{i}"] = np.array([f.score for f in faces], np.float32)
        out[f"lm{i}"] = np.array([f.landmarks for f in faces], np.float32).reshape(-1, 10)
        out[f"anchor{i}"] = np.array([f.anchor for f in faces], np.int32)
    np.savez_compressed(os.path.join(HERE, "decode_nms_case.npz"), **out)


def align_case():
    import cv2
    sys.path.insert(0, os.path.dirname(HERE))
    from conftest import synth_landmarks
    rng = np.random.default_rng(77)
    img = cv2.GaussianBlur(rng.integers(0, 256, (240, 320, 3), dtype=np.uint8), (0, 0), 1.5)
    lms = synth_landmarks(rng, 6, 320, 240, outlier_frac=0.5)
    crops = []
    for lm in lms:
        fb = odet.FaceBox(0, 0, 10, 10, 0.9, lm)
        crops.append(orec.align_face(img, fb))
    np.savez_compressed(os.path.join(HERE, "align_case.npz"), image=img, landmarks=lms,
                        crops=np.stack(crops))


if __name__ == "__main__":
    ref_arith()
    decode_nms_case()
    align_case()
    print("golden fixtures written to", HERE)
