"""loadModel-from-file, end to end, against an INDEPENDENT ONNX engine.

Reference: FaceDetector::loadModel / FaceRecognizer::loadModel open an .onnx file and run it through a
third-party executor (src/face_detector.cpp:20-90,179-183, src/face_recognizer.cpp:21-91,279-283).
Here: a raw, unfolded synthetic export (Conv -> BatchNormalization -> PRelu / Relu / Sigmoid / Resize /
Add / Flatten / Gemm; tests/onnx_emit.py) is written to disk; the PRODUCT loads that file through
``fr_weights_create(path)`` (its own protobuf reader + BN folding) and runs on the GPU; the CHECKER is
cv2.dnn executing the same file on the CPU (oracle/dnn_engine.py) -- it shares no code, no weight
layout and no graph restatement with the product or with oracle/nets.py.

Bars (north_star): scores within 1e-5... boxes / landmarks within 1e-3 px, embedding cosine >= 0.999, same
0.6 decisions; every detection-list difference explained (tests/parity.py).  Both with the seeded
random-init weights and with trained-like statistics (heavy-tailed BN scales, PReLU slopes 0.01..0.9,
residual stream |x| > 1e2; oracle/weights.py)."""
import numpy as np
import pytest

import onnx_emit
import parity
from conftest import SEED
from oracle import detector as odet
from oracle import dnn_engine
from oracle import recognizer as orec
from oracle import weights as ow

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["seeded", "trained_like"])
def onnx_models(request, tmp_path_factory, capi):
    d = tmp_path_factory.mktemp("models_" + request.param)
    pd, pr = str(d / "det_500m.onnx"), str(d / "w600k_r50.onnx")
    if request.param == "seeded":
        wd, wr = ow.seeded(ow.MODEL_DET, SEED), ow.seeded(ow.MODEL_REC, SEED)
    else:
        wd, wr = ow.trained_like_det(7), ow.trained_like_rec(7)
    onnx_emit.emit_det_full(wd, pd, raw=True, bbox_scales=(0.9, 1.7, 3.1), shuffle_seed=3)
    onnx_emit.emit_rec_full(wr, pr, raw=True, shuffle_seed=4)
    det_w, rec_w = capi.Weights(capi.FR_MODEL_DET, pd), capi.Weights(capi.FR_MODEL_REC, pr)
    assert det_w.from_onnx and rec_w.from_onnx
    c = capi.Context(0, det_w, rec_w)
    yield request.param, c, dnn_engine.DnnDetector(pd), dnn_engine.DnnRecognizer(pr)
    c.close()


def test_scrfd_heads_from_onnx_file_vs_cv2_dnn(onnx_models):
    kind, c, dnn_det, _ = onnx_models
    rng = np.random.default_rng(51)
    x = ((rng.integers(0, 256, (2, 3, 640, 640)).astype(np.float32)) - 127.5) / 128
    got = c.scrfd_forward(x)
    ref = dnn_det.heads(x)
    errs = [float(np.abs(g - r).max()) for g, r in zip(got, ref)]
    print(kind, "head max abs err vs cv2.dnn:", errs)
    assert max(errs[:3]) < 2e-5, errs                                  # sigmoid scores
    px = [e * s for e, s in zip(errs[3:], (8, 16, 32, 8, 16, 32))]
    assert max(px) < 1e-3, px                                          # boxes / landmarks in pixels
    # the yardstick: how far are the two fp32 engines from the float64 evaluation of the same graph?
    import torch
    from oracle import nets
    w = ow.seeded(ow.MODEL_DET, SEED) if kind == "seeded" else ow.trained_like_det(7)
    with nets.reference_precision():
        ref64 = [h.numpy() for h in nets.scrfd_forward(w, torch.from_numpy(x.astype(np.float64)))]
    e_gpu = [float(np.abs(g - r).max()) for g, r in zip(got, ref64)]
    e_dnn = [float(np.abs(g - r).max()) for g, r in zip(ref, ref64)]
    print(kind, "err vs float64: gpu", e_gpu, "cv2.dnn", e_dnn)
    for k in range(9):   # fp32-grade: within 4x of the fp32 CPU engine's own rounding error (+ 1 ulp-ish floor)
        assert e_gpu[k] <= 4 * e_dnn[k] + 2e-6, (k, e_gpu[k], e_dnn[k])


def test_detect_from_onnx_file_vs_cv2_dnn(onnx_models):
    """fr_detect on a ctx built from the file == the reference's detect() with cv2.dnn as the engine."""
    kind, c, dnn_det, _ = onnx_models
    rng = np.random.default_rng(52)
    n_common = 0
    for shape in ((640, 640, 3), (640, 640, 3), (480, 640, 3), (700, 500, 3)):
        im = rng.integers(0, 256, shape, dtype=np.uint8)
        chw, scale = odet.preprocess(im)
        heads = [h[0] for h in dnn_det.heads(chw[None])]
        for thr in ((0.5, 0.3) if kind == "trained_like" else (0.5, 0.02)):
            got = c.detect(im, thr, 0.4, cap=4096)
            st = parity.assert_detections_explained(got, heads, scale, thr, 0.4)
            n_common += st["common"]
    assert n_common > 50, n_common


def test_embed_from_onnx_file_vs_cv2_dnn(onnx_models):
    """fr_embed_aligned_batch / fr_embed (through alignment) on the file-loaded ctx vs cv2.dnn."""
    import cv2
    from conftest import faces_from_landmarks, synth_landmarks
    from facerecognizeonnx_b200 import capi
    kind, c, _, dnn_rec = onnx_models
    rng = np.random.default_rng(53)
    crops = rng.integers(0, 256, (6, 112, 112, 3), dtype=np.uint8)
    crops[4] = cv2.GaussianBlur(crops[4], (0, 0), 2.0)
    crops[5] = 255
    emb = c.embed_aligned(crops)
    ref = orec.normalize_rows(dnn_rec.embed(np.stack([orec.preprocess(x) for x in crops])))
    cos = (emb * ref).sum(1)
    print(kind, "cosine vs cv2.dnn:", cos)
    assert cos.min() >= 0.999, cos
    for i in range(6):
        for j in range(i + 1, 6):
            s_ref = orec.compare_faces(ref[i], ref[j])
            if abs(float(s_ref) - 0.6) > 2e-3:
                assert orec.same_person(capi.compare(emb[i], emb[j])) == orec.same_person(s_ref)
    # extractFeature (src/face_recognizer.cpp:236-304): align on the GPU, cv2 + cv2.dnn on the CPU
    img = cv2.GaussianBlur(rng.integers(0, 256, (480, 640, 3), dtype=np.uint8), (0, 0), 1.2)
    lms = synth_landmarks(rng, 4, 640, 480, outlier_frac=0.0)
    faces = faces_from_landmarks(capi, lms)
    e2, valid = c.embed_faces([img], faces, [0] * 4)
    assert valid.all()
    for i in range(4):
        fb = odet.FaceBox(int(faces[i]["x"]), int(faces[i]["y"]), int(faces[i]["w"]), int(faces[i]["h"]), 0.9, lms[i])
        aligned = orec.align_face(img, fb)
        r = orec.normalize(dnn_rec.embed(orec.preprocess(aligned)[None])[0])
        assert float((e2[i] * r).sum()) >= 0.999
    # extractFeatureSimple (src/face_recognizer.cpp:152-234)
    r = orec.normalize(dnn_rec.embed(orec.preprocess(cv2.resize(img, (112, 112)))[None])[0])
    assert float((c.embed_simple(img) * r).sum()) >= 0.999
