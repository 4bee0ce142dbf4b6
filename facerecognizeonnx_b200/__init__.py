"""B200-native face pipeline: the hot path of cucibala/FaceRecognizeOnnx behind its own API.

The product is ``libfr_b200.so`` (hand-written sm_100a CUDA behind the C ABI in
``include/fr_capi.h``).  ``capi`` binds it with ctypes; ``api`` mirrors the reference's
``FaceDetector`` / ``FaceRecognizer`` classes (src/face_detector.h, src/face_recognizer.h).
Importing this package never falls back to a CPU implementation.
"""
from . import capi  # noqa: F401
from .api import FaceBox, FaceDetector, FaceRecognizer  # noqa: F401

__all__ = ["capi", "FaceBox", "FaceDetector", "FaceRecognizer"]
