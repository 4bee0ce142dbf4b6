"""CPU oracle for the face-pipeline hot path of cucibala/FaceRecognizeOnnx.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and there only as
the checker / the CPU arm that is timed beside the GPU path.

What it restates (file:line are relative to the reference checkout):

* ``src/face_detector.cpp:92-137``   letterbox preprocess          -> ``detector.preprocess``
* ``src/face_detector.cpp:224-338``  postprocess (thr, /scale, int rect) -> ``detector.postprocess``
* ``src/face_detector.cpp:340-354``  integer IoU                   -> ``detector.iou``
* ``src/face_detector.cpp:356-384``  greedy NMS                    -> ``detector.nms``
* ``src/face_recognizer.cpp:93-133`` alignFace                     -> ``recognizer.align_face``
* ``src/face_recognizer.cpp:135-150`` preprocess                   -> ``recognizer.preprocess``
* ``src/face_recognizer.cpp:306-318`` normalize                    -> ``recognizer.normalize``
* ``src/face_recognizer.cpp:320-334`` compareFaces                 -> ``recognizer.compare_faces``

The dense arithmetic of the reference lives in two un-vendored third-party
libraries (ONNX Runtime >=1.12, README pins v1.16.3; OpenCV >=4.x).  Neither
is installable here, so:

* OpenCV calls are made through the ``cv2`` 4.13.0 wheel that *is* in the
  image (same algorithms), and additionally restated as integer recipes in
  ``cv_recipes.py`` which are pinned bit-for-bit against ``cv2`` by
  ``tests/test_oracle_cv.py``.
* ``Ort::Session::Run`` is restated as torch-CPU fp32 modules (``nets.py``)
  for the two fixed architectures (SCRFD det_500m, ArcFace IResNet-50).

PARITY PIN STATUS: the reference ships no tests, fixtures or golden vectors
for this path and cannot be compiled here
(OpenCV C++ / ONNX Runtime absent).  The reference-owned arithmetic
(threshold, /scale, int rects, integer IoU, greedy NMS, L2 norm, (dot+1)/2)
is pinned by hand-derived known-answer cases in ``tests/golden/``; the OpenCV
pieces are pinned against ``cv2`` itself; the two networks are
"parity unpinned" against ONNX Runtime (no .onnx file, no ORT) and are
checked only as torch-fp32 vs the CUDA path on shared weights.
"""
