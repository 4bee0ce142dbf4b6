"""Host-side checks of the C ABI that need no GPU: the library loads, exports every symbol
include/fr_capi.h declares, the weight store matches the oracle's architecture tables, and the
pure-host entry points behave like the reference."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from oracle import nets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(capi):
    hdr = open(os.path.join(ROOT, "include", "fr_capi.h")).read()
    declared = sorted(set(re.findall(r"FR_API\s+[\w\s\*]+?\b(fr_\w+)\s*\(", hdr)))
    assert len(declared) >= 40
    assert sorted(capi.SYMBOLS) == declared
    L = capi.lib()
    for name in declared:
        assert hasattr(L, name), name


def test_face_record_layout_matches_facebox(capi):
    # struct FaceBox { cv::Rect box; float score; cv::Point2f landmarks[5]; } = 60 bytes
    assert capi.FACE_DTYPE.itemsize == 60
    assert capi.FACE_DTYPE.fields["score"][1] == 16 and capi.FACE_DTYPE.fields["lm"][1] == 20


def test_weight_specs_match_oracle(capi, det_weights, rec_weights):
    assert det_weights.specs() == nets.det_tensor_specs()
    assert rec_weights.specs() == nets.rec_tensor_specs()
    n_rec = sum(int(np.prod(s)) for _, s in rec_weights.specs())
    n_det = sum(int(np.prod(s)) for _, s in det_weights.specs())
    assert 43.0e6 < n_rec < 44.0e6     # ~43.59 M (models/README.md:50-51: ~166 MB fp32)
    assert 0.55e6 < n_det < 0.70e6     # ~0.63 M (~2.5 MB)


def test_weights_are_seed_deterministic(capi):
    a = capi.Weights(capi.FR_MODEL_DET, None, 5).to_dict()
    b = capi.Weights(capi.FR_MODEL_DET, None, 5).to_dict()
    c = capi.Weights(capi.FR_MODEL_DET, None, 6).to_dict()
    assert all(np.array_equal(a[k], b[k]) for k in a)
    assert any(not np.array_equal(a[k], c[k]) for k in a)
    assert not capi.Weights(capi.FR_MODEL_DET, None, 5).from_onnx


def test_missing_or_bad_model_file_fails_like_loadmodel(capi, tmp_path):
    p = tmp_path / "garbage.onnx"
    p.write_bytes(b"not an onnx file")
    with pytest.raises(capi.FrError) as e:
        capi.Weights(capi.FR_MODEL_REC, str(p))
    assert e.value.code == capi.FR_ERR_MODEL


def test_tensor_set_roundtrip(capi):
    w = capi.Weights(capi.FR_MODEL_DET, None, 1)
    v = np.arange(16, dtype=np.float32)
    w.set("stem.b", v)
    assert np.array_equal(w.to_dict()["stem.b"], v)


def test_compare_host_matches_reference_semantics(capi):
    G = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_arith.json")))
    for c in G["compare"]:
        a, b = np.array(c["a"], np.float32), np.array(c["b"], np.float32)
        assert abs(capi.compare(a, b) - c["sim"]) < 1e-6
    from oracle import recognizer as orec
    rng = np.random.default_rng(0)
    for _ in range(20):
        a = orec.normalize(rng.normal(size=512).astype(np.float32))
        b = orec.normalize(rng.normal(size=512).astype(np.float32))
        assert np.float32(capi.compare(a, b)) == orec.compare_faces(a, b)  # same sequential fp32 order


def test_create_without_gpu_fails_loudly(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.FrError):
        capi.Context(0, None, None)
