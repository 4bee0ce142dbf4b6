"""The C++ ONNX reader (reference src/face_detector.cpp:20-90,
src/face_recognizer.cpp:21-91) against synthetic .onnx files with the graph structure of the
buffalo exports (tests/onnx_emit.py): raw (Conv -> BatchNormalization unfolded) and exporter-fused
forms, shuffled node order, the mmdet Scale nodes, wrong architectures, and malformed / truncated
files.  Host only, no GPU."""
import numpy as np
import pytest

import onnx_emit


def _maxrel(a, b):
    return max(float(np.abs(a[k] - b[k]).max() / (np.abs(b[k]).max() + 1e-12)) for k in b)


@pytest.mark.parametrize("raw,shuffle", [(False, None), (True, None), (True, 5)])
def test_rec_onnx_roundtrip(capi, rec_wdict, tmp_path, raw, shuffle):
    p = str(tmp_path / "w600k_r50.onnx")
    onnx_emit.emit_rec_full(rec_wdict, p, raw=raw, shuffle_seed=shuffle)
    w = capi.Weights(capi.FR_MODEL_REC, p)
    assert w.from_onnx
    got = w.to_dict()
    assert list(got) == list(rec_wdict)
    if not raw:
        exact = [k for k in got if "bn" not in k and not k.startswith("feat.")]
        assert all(np.array_equal(got[k], rec_wdict[k]) for k in exact)
    assert _maxrel(got, rec_wdict) < (3e-5 if raw else 3e-6)   # fp32 rounding of the raw BN parameters only


@pytest.mark.parametrize("raw,shuffle", [(False, None), (True, None), (True, 9)])
def test_det_onnx_roundtrip_with_bbox_scale(capi, det_wdict, tmp_path, raw, shuffle):
    p = str(tmp_path / "det_500m.onnx")
    onnx_emit.emit_det_full(det_wdict, p, raw=raw, bbox_scales=(0.9, 1.7, 3.1), shuffle_seed=shuffle)
    got = capi.Weights(capi.FR_MODEL_DET, p).to_dict()
    assert list(got) == list(det_wdict)
    assert _maxrel(got, det_wdict) < (3e-5 if raw else 3e-6)
    if not raw:
        p2 = str(tmp_path / "det_noscale.onnx")
        onnx_emit.emit_det_full(det_wdict, p2, raw=False)
        got2 = capi.Weights(capi.FR_MODEL_DET, p2).to_dict()
        assert all(np.array_equal(got2[k], det_wdict[k]) for k in got2)


def test_same_shape_layers_are_bound_by_edges_not_by_file_order(capi, det_wdict, tmp_path):
    """fpn0..2 / pafpn0..1 / down0..1 / the three head towers have identical weight shapes: a
    reader that walks nodes in file order permutes them silently when an exporter reorders nodes."""
    p = str(tmp_path / "det.onnx")
    for seed, topo in ((1, True), (2, True), (3, False), (4, False)):
        onnx_emit.emit_det_full(det_wdict, p, raw=False, shuffle_seed=seed, topological=topo)
        got = capi.Weights(capi.FR_MODEL_DET, p).to_dict()
        for k in ("fpn0.w", "fpn1.w", "fpn2.w", "pafpn0.w", "pafpn1.w", "down0.w", "down1.w",
                  "h0.t1.pw.w", "h1.t1.pw.w", "h2.t1.pw.w", "h0.kps.b", "h2.kps.b"):
            assert np.array_equal(got[k], det_wdict[k]), k


def test_wrong_architecture_fails_loudly(capi, det_wdict, tmp_path):
    p = str(tmp_path / "det_500m.onnx")
    onnx_emit.emit_det_full(det_wdict, p)
    with pytest.raises(capi.FrError) as e:
        capi.Weights(capi.FR_MODEL_REC, p)          # a detector file is not an IResNet-50
    assert e.value.code == capi.FR_ERR_MODEL and "IResNet-50" in str(e.value)
    bad = dict(det_wdict)
    bad["s1.0.pw.w"] = np.zeros((72, 41, 1, 1), np.float32)
    p2 = str(tmp_path / "bad.onnx")
    onnx_emit.emit_det_full(bad, p2, raw=False)
    with pytest.raises(capi.FrError) as e:
        capi.Weights(capi.FR_MODEL_DET, p2)
    assert "s1.0.pw" in str(e.value)
    with pytest.raises(capi.FrError):
        capi.Weights(capi.FR_MODEL_DET, str(tmp_path / "missing.onnx"))


def test_malformed_files_are_rejected_not_read_out_of_bounds(capi, det_wdict, tmp_path):
    """Truncations at every kind of boundary, flipped length bytes and absurd dims must end in
    FR_ERR_MODEL (the reader checks the remaining bytes before every fixed-width read / skip)."""
    p = str(tmp_path / "det.onnx")
    onnx_emit.emit_det_full(det_wdict, p, raw=True)
    blob = open(p, "rb").read()
    rng = np.random.default_rng(0)
    cuts = [0, 1, 2, 3, 7, len(blob) // 2, len(blob) - 1, len(blob) - 3] + [int(c) for c in rng.integers(4, len(blob), 40)]
    q = str(tmp_path / "cut.onnx")
    for c in cuts:
        open(q, "wb").write(blob[:c])
        with pytest.raises(capi.FrError) as e:
            capi.Weights(capi.FR_MODEL_DET, q)
        assert e.value.code == capi.FR_ERR_MODEL
    for _ in range(60):                               # random byte corruption: load or fail, never crash
        b = bytearray(blob)
        for pos in rng.integers(0, min(len(b), 200000), 8):
            b[int(pos)] = int(rng.integers(0, 256))
        open(q, "wb").write(bytes(b))
        try:
            capi.Weights(capi.FR_MODEL_DET, q)
        except capi.FrError as e:
            assert e.code == capi.FR_ERR_MODEL
    # a tensor whose dims multiply past any sane size
    huge = onnx_emit._key(1, 0) + onnx_emit._varint(1 << 40) + onnx_emit._key(1, 0) + onnx_emit._varint(1 << 40)
    huge += onnx_emit._key(2, 0) + onnx_emit._varint(1) + onnx_emit._ld(8, b"w")
    g = onnx_emit._ld(5, huge)
    open(q, "wb").write(onnx_emit._key(1, 0) + onnx_emit._varint(6) + onnx_emit._ld(7, g))
    with pytest.raises(capi.FrError):
        capi.Weights(capi.FR_MODEL_DET, q)
