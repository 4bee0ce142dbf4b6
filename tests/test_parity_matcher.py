"""CPU self-test of tests/parity.py: the explained-difference matcher accepts what a second fp32 engine
produces (oracle heads + 1e-6 noise, hundreds of candidates per frame) and rejects real errors."""
import numpy as np
import pytest
import torch

import parity
from oracle import detector as odet
from oracle import nets
from oracle import weights as ow

FACE_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("w", "<i4"), ("h", "<i4"), ("score", "<f4"), ("lm", "<f4", (10,))])


def _records(boxes):
    out = np.zeros(len(boxes), FACE_DTYPE)
    for r, b in zip(out, boxes):
        r["x"], r["y"], r["w"], r["h"], r["score"] = b.x, b.y, b.w, b.h, b.score
        r["lm"] = b.landmarks.reshape(10)
    return out


@pytest.fixture(scope="module")
def frames_heads():
    w = ow.trained_like_det(11)
    rng = np.random.default_rng(12)
    out = []
    for k in range(4):
        im = rng.integers(0, 256, (640, 640, 3) if k < 3 else (480, 600, 3), dtype=np.uint8)
        chw, scale = odet.preprocess(im)
        heads = [h[0].numpy() for h in nets.scrfd_forward(w, torch.from_numpy(chw[None]))]
        out.append((heads, scale))
    return out


@pytest.mark.parametrize("thr", [0.5, 0.3])
def test_matcher_accepts_a_second_fp32_engine(frames_heads, thr):
    rng = np.random.default_rng(13)
    total = {"common": 0, "missing": 0, "extra": 0}
    for heads, scale in frames_heads:
        for rep in range(3):
            noisy = [h + (rng.normal(0, 1e-6, h.shape) * np.maximum(1.0, np.abs(h))).astype(np.float32) for h in heads]
            got = _records(odet.postprocess(odet.scrfd_decode(noisy), scale, thr, 0.4))
            st = parity.assert_detections_explained(got, heads, scale, thr, 0.4)
            for k in total:
                total[k] += st[k]
    assert total["common"] > 200, total        # the frames really have detections to compare


def test_matcher_rejects_real_errors(frames_heads):
    heads, scale = frames_heads[0]
    exp = odet.postprocess(odet.scrfd_decode(heads), scale, 0.5, 0.4)
    clear = [i for i, b in enumerate(exp) if b.score > 0.52]
    assert len(clear) > 5
    got = _records(exp)
    parity.assert_detections_explained(got, heads, scale, 0.5, 0.4)
    with pytest.raises(AssertionError):                       # a clearly kept detection dropped
        parity.assert_detections_explained(np.delete(got, clear[2]), heads, scale, 0.5, 0.4)
    bad = got.copy()
    bad[clear[1]]["x"] += 3                                   # rect off by 3 px
    with pytest.raises(AssertionError):
        parity.assert_detections_explained(bad, heads, scale, 0.5, 0.4)
    bad = got.copy()
    bad[clear[1]]["lm"][4] += 0.01                            # landmark off by 1e-2 px
    with pytest.raises(AssertionError):
        parity.assert_detections_explained(bad, heads, scale, 0.5, 0.4)
    # a suppressed candidate resurrected (NMS skipped for it)
    all_c = odet.postprocess(odet.scrfd_decode(heads), scale, 0.5, 1.1)
    kept = {b.anchor for b in exp}
    sup = [b for b in all_c if b.anchor not in kept and b.score > 0.51]
    if sup:
        extra = sorted(exp + [sup[0]], key=lambda b: -b.score)
        with pytest.raises(AssertionError):
            parity.assert_detections_explained(_records(extra), heads, scale, 0.5, 0.4)
