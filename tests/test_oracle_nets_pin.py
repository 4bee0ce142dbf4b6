"""Pins oracle/nets.py (torch-CPU restatement of the two Ort::Session::Run calls, reference
src/face_detector.cpp:179-183, src/face_recognizer.cpp:279-283) against an INDEPENDENT ONNX engine:
cv2.dnn executing the raw .onnx graph (Conv -> BatchNormalization -> PRelu / Relu / Sigmoid / Resize /
Add / Flatten / Gemm, nothing folded) written by tests/onnx_emit.py.  Also pins the pure-NumPy twin of
the product's seeded weight generator (oracle/weights.py) to the C ABI's output.  CPU only."""
import numpy as np
import pytest
import torch

import onnx_emit
from conftest import SEED
from oracle import dnn_engine, nets
from oracle import weights as ow


def test_numpy_seeded_weights_equal_c_abi(capi):
    for model in (capi.FR_MODEL_DET, capi.FR_MODEL_REC):
        for seed in (SEED, 0x1234567890ABCDEF):
            a = capi.Weights(model, None, seed).to_dict()
            b = ow.seeded(model, seed)
            assert list(a) == list(b)
            assert all(np.array_equal(a[k], b[k]) for k in a)


@pytest.mark.parametrize("kind", ["seeded", "trained_like"])
def test_scrfd_oracle_equals_cv2_dnn_on_raw_graph(tmp_path, kind):
    w = ow.seeded(ow.MODEL_DET, SEED) if kind == "seeded" else ow.trained_like_det(3)
    p = str(tmp_path / "det_500m.onnx")
    onnx_emit.emit_det_full(w, p, raw=True, bbox_scales=(0.9, 1.7, 3.1))
    rng = np.random.default_rng(5)
    x = ((rng.integers(0, 256, (2, 3, 640, 640)).astype(np.float32)) - 127.5) / 128
    got = dnn_engine.DnnDetector(p).heads(x)
    ref = nets.scrfd_forward(w, torch.from_numpy(x))
    for k, (g, r) in enumerate(zip(got, ref)):
        r = r.numpy()
        assert g.shape == r.shape
        err = float(np.abs(g - r).max())
        scale = max(1.0, float(np.abs(r).max()))
        assert err < 1e-5 * scale, (k, err, scale)      # two fp32 engines: summation order only
    if kind == "trained_like":
        n_pos = sum(int((r.numpy() > 0.5).sum()) for r in ref[:3])
        assert 20 <= n_pos <= 2000, n_pos                # decode / NMS will see real candidate counts


@pytest.mark.parametrize("kind", ["seeded", "trained_like"])
def test_iresnet_oracle_equals_cv2_dnn_on_raw_graph(tmp_path, kind):
    w = ow.seeded(ow.MODEL_REC, SEED) if kind == "seeded" else ow.trained_like_rec(3)
    p = str(tmp_path / "w600k_r50.onnx")
    onnx_emit.emit_rec_full(w, p, raw=True)
    rng = np.random.default_rng(6)
    x = ((rng.integers(0, 256, (2, 3, 112, 112)).astype(np.float32)) - 127.5) / 128
    got = dnn_engine.DnnRecognizer(p).embed(x)
    ref, taps = nets.iresnet50_forward(w, torch.from_numpy(x), return_taps=True)
    ref = ref.numpy()
    err = float(np.abs(got - ref).max() / np.abs(ref).max())
    assert err < 1e-5 if kind == "seeded" else err < 1e-4, err
    cos = (got * ref).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(ref, axis=1))
    assert cos.min() > 1 - 1e-6
    if kind == "trained_like":
        # the statistics the generator promises: big residual stream, wide slopes
        peak = max(float(t.abs().max()) for k, t in taps.items() if not k.endswith(".h"))
        assert peak > 1e2, peak
        slopes = np.concatenate([v for k, v in w.items() if k.endswith("prelu")])
        assert slopes.min() <= 0.011 and slopes.max() >= 0.89
        sc = np.concatenate([np.abs(v) for k, v in w.items() if k.endswith("bn1.scale")])
        assert np.isfinite(ref).all()
