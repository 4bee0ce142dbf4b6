"""Pins the oracle's restatement of the reference-owned arithmetic against the hand-derived
known-answer cases in tests/golden/ref_arith.json (CPU only)."""
import json
import math
import os

import numpy as np

from oracle import detector as odet
from oracle import gallery as ogal
from oracle import recognizer as orec

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_arith.json")))


def _fb(b, score=0.5, anchor=-1):
    return odet.FaceBox(b[0], b[1], b[2], b[3], score, anchor=anchor)


def test_iou_known_answers():
    for c in G["iou"]:
        v = odet.iou(_fb(c["a"]), _fb(c["b"]))
        if c["den"] == 0:
            assert math.isnan(float(v))
            assert not (v > np.float32(0.0))
        else:
            assert v == np.float32(c["num"]) / np.float32(c["den"])


def test_rect_truncation_known_answers():
    for c in G["rect"]:
        row = np.zeros((1, 15), np.float32)
        row[0, :4] = c["row"]
        row[0, 4] = 0.9
        faces = odet.postprocess(row, c["scale"], 0.5, 0.4)
        assert len(faces) == 1
        f = faces[0]
        assert [f.x, f.y, f.w, f.h] == c["rect"]


def test_nms_known_answers():
    for c in G["nms"]:
        boxes = [_fb(b[:4], b[4], anchor=i) for i, b in enumerate(c["boxes"])]
        kept = odet.nms(boxes, c["thr"])
        assert [k.anchor for k in kept] == c["keep"]


def test_score_threshold_is_strict():
    row = np.zeros((2, 15), np.float32)
    row[:, 2:4] = 10
    row[0, 4] = 0.5
    row[1, 4] = np.nextafter(np.float32(0.5), np.float32(1))
    faces = odet.postprocess(row, 1.0, 0.5, 0.4)
    assert len(faces) == 1 and faces[0].anchor == 1


def test_compare_and_normalize_known_answers():
    for c in G["compare"]:
        assert abs(float(orec.compare_faces(np.array(c["a"], np.float32), np.array(c["b"], np.float32))) - c["sim"]) < 1e-6
    for c in G["normalize"]:
        assert np.allclose(orec.normalize(np.array(c["v"], np.float32)), np.array(c["out"], np.float32), atol=1e-7)
    for c in G["decision"]:
        assert orec.same_person(np.float32(c["sim"])) == c["same"]


def test_letterbox_geometry():
    # fp32 scale, truncated sizes (face_detector.cpp:101-106)
    assert odet.letterbox_geometry(640, 640)[1:] == (640, 640)
    assert odet.letterbox_geometry(720, 1280)[1:] == (640, 360)
    s, w, h = odet.letterbox_geometry(333, 517)
    assert (w, h) == (640, 412) and s == np.float32(640) / np.float32(517)


def test_preprocess_pad_value_and_layout():
    img = np.zeros((100, 200, 3), np.uint8)
    img[..., 0] = 10   # B
    img[..., 2] = 250  # R
    chw, scale = odet.preprocess(img)
    assert chw.shape == (3, 640, 640) and scale == np.float32(3.2)
    assert chw[0, 0, 0] == np.float32((250 - 127.5) / 128)  # plane 0 = R
    assert chw[2, 0, 0] == np.float32((10 - 127.5) / 128)
    assert chw[0, 639, 639] == np.float32(-0.99609375)      # zero padding, bottom-right
    assert odet.preprocess(np.zeros((0, 0, 3), np.uint8))[0] is None
    assert odet.preprocess(np.zeros((1, 10000, 3), np.uint8))[0] is None  # int(1*0.064) == 0


def test_decode_layout_and_values():
    heads = [np.zeros((n, c), np.float32) for c in (1, 4, 10) for n in (12800, 3200, 800)]
    # stride 16, cell (gy=3, gx=5), anchor 1 -> local index (3*40+5)*2+1 = 251
    heads[1][251, 0] = 0.9
    heads[4][251] = [1.0, 2.0, 3.0, 4.0]
    heads[7][251] = np.arange(10) * 0.5
    out = odet.scrfd_decode(heads)
    assert out.shape == (16800, 15)
    r = out[12800 + 251]
    assert list(r[:5]) == [80 - 16, 48 - 32, 80 + 48, 48 + 64, np.float32(0.9)]
    assert r[5] == 80 and r[6] == 48 + 8 and r[13] == 80 + 64 and r[14] == 48 + 72


def test_bf16_rounding_and_topk_ties():
    x = np.array([1.0, 1.00390625, 1.005859375], np.float32)  # exact, tie-to-even, above-half
    assert list(ogal.to_bf16_f32(x)) == [1.0, 1.0, 1.0078125]
    q = np.eye(4, dtype=np.float32)[:1]
    g = np.eye(4, dtype=np.float32)[[0, 0, 1, 0]]
    s, i = ogal.topk(q, g, 3)
    assert list(i[0]) == [0, 1, 3] and list(s[0]) == [1.0, 1.0, 1.0]  # ties -> lower index first
    ms, mi = ogal.merge_topk([s[:, :2], s[:, 2:]], [i[:, :2], i[:, 2:]], 3)
    assert list(mi[0]) == [0, 1, 3]
