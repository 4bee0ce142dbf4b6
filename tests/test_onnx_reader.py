"""The C++ ONNX initializer reader against synthetic .onnx files with the graph
structure of the buffalo exports (tests/onnx_emit.py).  Host only, no GPU."""
import numpy as np
import pytest

import onnx_emit


def _close(a, b, tol):
    return all(np.allclose(a[k], b[k], rtol=tol, atol=tol) for k in a)


@pytest.mark.parametrize("trans_b", [True, False])
def test_rec_onnx_roundtrip(capi, rec_wdict, tmp_path, trans_b):
    p = str(tmp_path / "w600k_r50.onnx")
    onnx_emit.emit_rec(rec_wdict, p, trans_b=trans_b)
    w = capi.Weights(capi.FR_MODEL_REC, p)
    assert w.from_onnx
    got = w.to_dict()
    assert set(got) == set(rec_wdict)
    exact = [k for k in got if "bn" not in k and not k.startswith("feat.")]
    assert all(np.array_equal(got[k], rec_wdict[k]) for k in exact)
    assert _close({k: got[k] for k in got if k not in exact}, rec_wdict, 2e-6)   # BN folding rounding


def test_det_onnx_roundtrip_with_bbox_scale(capi, det_wdict, tmp_path):
    p = str(tmp_path / "det_500m.onnx")
    onnx_emit.emit_det(det_wdict, p, bbox_scales=(0.9, 1.7, 3.1))
    got = capi.Weights(capi.FR_MODEL_DET, p).to_dict()
    assert _close(got, det_wdict, 1e-6)
    p2 = str(tmp_path / "det_noscale.onnx")
    onnx_emit.emit_det(det_wdict, p2, with_scale_nodes=False)
    got2 = capi.Weights(capi.FR_MODEL_DET, p2).to_dict()
    assert all(np.array_equal(got2[k], det_wdict[k]) for k in got2)


def test_wrong_architecture_fails_loudly(capi, det_wdict, tmp_path):
    p = str(tmp_path / "det_500m.onnx")
    onnx_emit.emit_det(det_wdict, p)
    with pytest.raises(capi.FrError) as e:
        capi.Weights(capi.FR_MODEL_REC, p)          # a detector file is not an IResNet-50
    assert e.value.code == capi.FR_ERR_MODEL and "IResNet-50" in str(e.value)
    bad = dict(det_wdict)
    bad["s1.0.pw.w"] = np.zeros((72, 41, 1, 1), np.float32)
    p2 = str(tmp_path / "bad.onnx")
    onnx_emit.emit_det(bad, p2)
    with pytest.raises(capi.FrError) as e:
        capi.Weights(capi.FR_MODEL_DET, p2)
    assert "s1.0.pw" in str(e.value)
    with pytest.raises(capi.FrError):
        capi.Weights(capi.FR_MODEL_DET, str(tmp_path / "missing.onnx"))
