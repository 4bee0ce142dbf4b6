"""Second, independent ONNX CPU engine: OpenCV's dnn module executing the .onnx graph itself.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The reference runs its two graphs through ``Ort::Session::Run`` (src/face_detector.cpp:179-183,
src/face_recognizer.cpp:279-283).  ONNX Runtime is not installable here, but ``cv2.dnn`` is in the
image and is an ONNX executor written by other people: it parses the protobuf, builds its own layer
graph (Conv / BatchNorm / PReLU / Eltwise / Resize / Permute / Reshape / InnerProduct) and fuses what it
likes.  Two restatements by one author (oracle/nets.py and the CUDA kernels) agreeing with each other
proves little; both agreeing with an engine that only ever saw the ``.onnx`` FILE pins
  * oracle/nets.py's reading of the graph (tests/test_oracle_nets_pin.py, CPU), and
  * the product's whole ``loadModel(path)`` -> GPU forward path (tests/test_gpu_onnx.py).
Winograd convolution is switched off: it trades ~1e-5 of accuracy for speed, and this is the checker.
"""
from __future__ import annotations

from typing import List

import numpy as np

DET_OUT_NAMES = ["score_8", "score_16", "score_32", "bbox_8", "bbox_16", "bbox_32", "kps_8", "kps_16", "kps_32"]
_HEAD_C = (1, 1, 1, 4, 4, 4, 10, 10, 10)


class DnnNet:
    def __init__(self, onnx_path: str):
        import cv2
        self.net = cv2.dnn.readNetFromONNX(onnx_path)
        self.net.enableWinograd(False)
        self.net.setPreferableBackend(cv2.dnn.DNN_BACKEND_OPENCV)
        self.net.setPreferableTarget(cv2.dnn.DNN_TARGET_CPU)
        self.out_names = list(self.net.getUnconnectedOutLayersNames())

    def run(self, x: np.ndarray, names=None) -> List[np.ndarray]:
        self.net.setInput(np.ascontiguousarray(x, np.float32))
        outs = self.net.forward(names if names is not None else self.out_names)
        return [np.array(o, np.float32) for o in outs]


class DnnRecognizer(DnnNet):
    """w600k_r50-shaped file: [B,3,112,112] RGB in [-1,1] -> [B,512] (not normalised), one frame per
    forward like the reference (src/face_recognizer.cpp:262-283)."""

    def embed(self, chw: np.ndarray) -> np.ndarray:
        chw = np.ascontiguousarray(chw, np.float32)
        return np.concatenate([self.run(chw[i:i + 1])[0].reshape(1, -1) for i in range(chw.shape[0])])


class DnnDetector(DnnNet):
    """det_500m-shaped file: [1,3,640,640] -> the nine head tensors [A_s, c] of the buffalo export
    (models/README.md:12), returned with a leading batch axis in oracle/nets.py's order."""

    def heads(self, chw: np.ndarray) -> List[np.ndarray]:
        chw = np.ascontiguousarray(chw, np.float32)
        per = []
        for i in range(chw.shape[0]):
            outs = self.run(chw[i:i + 1], DET_OUT_NAMES)
            per.append([o.reshape(-1, c) for o, c in zip(outs, _HEAD_C)])
        return [np.stack([p[k] for p in per]) for k in range(9)]
