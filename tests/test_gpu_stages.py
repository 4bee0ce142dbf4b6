"""GPU parity tests for the bandwidth / integer stages, through the C ABI, against the oracle.

Bars: K1 preprocess, cv::resize, cv::warpAffine (given M), decode+NMS kept set: BIT-EXACT.
Alignment estimate: |dM| <= 1e-6 vs cv2 (cv2's own LM refine stops ~2e-9 from the LS optimum).
"""
import os

import cv2
import numpy as np
import pytest

from conftest import faces_from_landmarks, synth_landmarks
from oracle import cv_recipes as R
from oracle import detector as odet
from oracle import recognizer as orec

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _img(rng, h, w):
    return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)


def test_k1_preprocess_bit_exact(ctx):
    rng = np.random.default_rng(0)
    imgs = [_img(rng, 640, 640), _img(rng, 720, 1280), _img(rng, 333, 517), _img(rng, 480, 640),
            _img(rng, 300, 200), _img(rng, 64, 48), _img(rng, 1080, 1920), _img(rng, 640, 360), _img(rng, 7, 1000)]
    out, scale = ctx.det_preprocess(imgs)
    for i, im in enumerate(imgs):
        ref, s = odet.preprocess(im)
        assert scale[i] == s
        assert np.array_equal(out[i], ref), f"image {i} {im.shape}: {np.abs(out[i]-ref).max()}"


def test_k1_preprocess_strided_rows(ctx):
    rng = np.random.default_rng(1)
    big = _img(rng, 500, 700)
    roi = big[10:410, 20:620]  # non-contiguous rows (cv::Mat ROI), step = 2100
    out, _ = ctx.det_preprocess([roi])
    ref, _ = odet.preprocess(np.ascontiguousarray(roi))
    assert np.array_equal(out[0], ref)


def test_k1_rejects_bad_images(ctx, capi):
    with pytest.raises(capi.FrError) as e:
        ctx.det_preprocess([np.zeros((1, 10000, 3), np.uint8)])  # int(1*0.064) == 0 (face_detector.cpp:109-113)
    assert e.value.code == capi.FR_ERR_INVALID_ARG


@pytest.mark.parametrize("sw,sh,nw,nh", [(1280, 720, 640, 360), (517, 333, 640, 412), (200, 300, 112, 112),
                                          (37, 91, 112, 112), (112, 112, 112, 112), (48, 64, 480, 640)])
def test_resize_bit_exact(ctx, sw, sh, nw, nh):
    rng = np.random.default_rng(sw + sh)
    im = _img(rng, sh, sw)
    assert np.array_equal(ctx.resize_linear(im, nw, nh), cv2.resize(im, (nw, nh)))


def test_k5_warp_affine_bit_exact(ctx):
    rng = np.random.default_rng(3)
    im = _img(rng, 480, 640)
    for _ in range(25):
        s, th = rng.uniform(0.3, 1.5), rng.uniform(-0.5, 0.5)
        M = np.array([[s * np.cos(th), -s * np.sin(th), rng.uniform(-150, 150)],
                      [s * np.sin(th), s * np.cos(th), rng.uniform(-150, 150)]])
        got = ctx.warp_affine(im, M)
        assert np.array_equal(got, cv2.warpAffine(im, M, (112, 112)))
        assert np.array_equal(got, R.warp_affine_u8(im, M))


def test_k5_estimate_matches_cv2(ctx):
    rng = np.random.default_rng(5)
    lms = synth_landmarks(rng, 500, outlier_frac=0.5)
    lms[7] = 0  # degenerate: all landmarks identical -> empty M -> fallback
    M, ok = ctx.estimate_alignment(lms)
    worst = 0.0
    for i, lm in enumerate(lms):
        Mc, _ = cv2.estimateAffinePartial2D(lm, orec.TEMPLATE)
        assert (Mc is not None) == bool(ok[i]), i
        if Mc is not None:
            worst = max(worst, float(np.abs(Mc - M[i]).max()))
    assert not ok[7]
    assert worst < 1e-6, worst


def test_k5_align_faces_vs_cv2(ctx, capi):
    g = np.load(os.path.join(GOLD, "align_case.npz"))
    img, lms, gold = g["image"], g["landmarks"], g["crops"]
    crops, valid = ctx.align_faces(img, faces_from_landmarks(capi, lms))
    assert valid.all()
    diff = np.abs(crops.astype(int) - gold.astype(int))
    # M agrees with cv2 to ~1e-9; a fixed-point coordinate can still flip at a rounding
    # boundary, so allow a vanishing fraction of 1-level differences.
    assert diff.max() <= 2 and (diff > 0).mean() < 1e-3, (diff.max(), (diff > 0).mean())


def test_k5_fallback_crop_and_invalid(ctx, capi):
    rng = np.random.default_rng(9)
    img = _img(rng, 200, 300)
    f = np.zeros(3, capi.FACE_DTYPE)
    f[0]["x"], f[0]["y"], f[0]["w"], f[0]["h"] = 250, 150, 100, 100   # clipped by the image
    f[1]["x"], f[1]["y"], f[1]["w"], f[1]["h"] = 400, 400, 10, 10     # fully outside -> empty
    f[2]["x"], f[2]["y"], f[2]["w"], f[2]["h"] = -20, -10, 60, 50     # clipped at the origin
    crops, valid = ctx.align_faces(img, f)  # all-zero landmarks -> estimateAffinePartial2D empty
    assert list(valid) == [1, 0, 1]
    assert np.array_equal(crops[0], cv2.resize(np.ascontiguousarray(img[150:200, 250:300]), (112, 112)))
    assert np.array_equal(crops[2], cv2.resize(np.ascontiguousarray(img[0:40, 0:40]), (112, 112)))
    for i in (0, 2):
        fb = odet.FaceBox(int(f[i]["x"]), int(f[i]["y"]), int(f[i]["w"]), int(f[i]["h"]), 0.9, np.zeros((5, 2), np.float32))
        assert np.array_equal(crops[i], orec.align_face(img, fb))


def _check_faces(got, exp_faces):
    assert len(got) == len(exp_faces)
    for r, e in zip(got, exp_faces):
        assert (int(r["x"]), int(r["y"]), int(r["w"]), int(r["h"])) == (e.x, e.y, e.w, e.h)
        assert np.float32(r["score"]) == np.float32(e.score)
        assert np.array_equal(np.array(r["lm"], np.float32), e.landmarks.reshape(10))


def test_k3_k4_decode_nms_golden_bit_exact(ctx):
    g = np.load(os.path.join(GOLD, "decode_nms_case.npz"))
    heads = [g[f"head{i}"] for i in range(9)]
    got = ctx.scrfd_decode_nms(heads, g["scales"], 0.5, 0.4, cap=512)
    for i in range(2):
        assert np.array_equal(np.stack([got[i]["x"], got[i]["y"], got[i]["w"], got[i]["h"]], 1).reshape(-1, 4), g[f"rect{i}"])
        assert np.array_equal(got[i]["score"], g[f"score{i}"])
        assert np.array_equal(got[i]["lm"].reshape(-1, 10), g[f"lm{i}"])
        exp = odet.postprocess(odet.scrfd_decode([h[i] for h in heads]), g["scales"][i], 0.5, 0.4)
        _check_faces(got[i], exp)


def _random_heads(rng, n, mean_logit, box_mean=1.8):
    heads = []
    for k, c in enumerate((1, 4, 10)):
        for ns in (12800, 3200, 800):
            if k == 0:
                a = 1.0 / (1.0 + np.exp(-rng.normal(mean_logit, 1.2, (n, ns, c))))
            elif k == 1:
                a = rng.normal(box_mean, 0.8, (n, ns, c))
            else:
                a = rng.normal(0.0, 1.0, (n, ns, c))
            heads.append(a.astype(np.float32))
    return heads


@pytest.mark.parametrize("mean_logit,thr,nms_thr", [(-3.5, 0.5, 0.4), (-2.0, 0.5, 0.4), (-3.0, 0.3, 0.1), (-6.0, 0.5, 0.4)])
def test_k3_k4_decode_nms_random(ctx, mean_logit, thr, nms_thr):
    rng = np.random.default_rng(int(-mean_logit * 10))
    n = 3
    heads = _random_heads(rng, n, mean_logit)
    scales = np.array([1.0, 0.5, 1.7777778], np.float32)
    got = ctx.scrfd_decode_nms(heads, scales, thr, nms_thr, cap=4096)
    for i in range(n):
        exp = odet.postprocess(odet.scrfd_decode([h[i] for h in heads]), scales[i], thr, nms_thr)
        _check_faces(got[i], exp)


def test_k3_k4_many_candidates_spill_path(ctx):
    """> 4096 candidates in one frame exercises the global-memory sort path; ties, zero-area
    and inverted boxes exercise the integer semantics."""
    rng = np.random.default_rng(11)
    heads = _random_heads(rng, 1, 0.3, box_mean=0.6)  # ~60 % of anchors pass 0.5
    heads[0][0, ::3] = 0.8125                           # massive exact ties
    heads[3][0, ::7] = 0.0                              # zero-size boxes -> NaN IoU
    heads[3][0, ::11] = -1.0                            # inverted boxes
    got = ctx.scrfd_decode_nms(heads, np.array([1.0], np.float32), 0.5, 0.4, cap=16800)
    exp = odet.postprocess(odet.scrfd_decode([h[0] for h in heads]), 1.0, 0.5, 0.4)
    assert len(exp) > 100
    _check_faces(got[0], exp)


@pytest.mark.parametrize("k", [1, 2, 31, 32, 33, 64, 511, 512, 513, 1000])
def test_k4_nms_candidate_count_boundaries(ctx, k):
    """Exactly k candidates per frame around the limits of the two NMS algorithms (bit matrix up to 512
    candidates, 32-candidate blocks above; warp / word boundaries at 32): clustered boxes so that chains of
    suppressions cross word boundaries, exact score ties, zero-area and inverted boxes."""
    rng = np.random.default_rng(100 + k)
    heads = _random_heads(rng, 2, -30.0)                 # nothing passes on its own
    for img in range(2):
        picks = rng.choice(12800 + 3200 + 800, size=k, replace=False)
        # neighbouring anchors overlap heavily: take half of the picks as runs of adjacent anchors
        run = min(k // 2, 12000)
        picks[:run] = 2000 + np.arange(run) + img
        picks = np.unique(picks)
        while len(picks) < k:                            # refill after de-duplication
            extra = rng.choice(16800, size=k - len(picks), replace=False)
            picks = np.unique(np.concatenate([picks, extra]))
        picks = picks[:k]
        sc = rng.uniform(0.55, 0.95, size=k).astype(np.float32)
        sc[::5] = 0.75                                   # ties
        for a, v in zip(picks, sc):
            s, local = (0, a) if a < 12800 else ((1, a - 12800) if a < 16000 else (2, a - 16000))
            heads[s][img, local, 0] = v
            if a % 13 == 0:
                heads[3 + s][img, local] = 0.0           # zero-area box: IoU 0/0 -> NaN -> never suppresses
            if a % 17 == 0:
                heads[3 + s][img, local] = -0.5          # inverted box
    scales = np.array([1.0, 0.75], np.float32)
    got = ctx.scrfd_decode_nms(heads, scales, 0.5, 0.4, cap=4096)
    for img in range(2):
        exp = odet.postprocess(odet.scrfd_decode([h[img] for h in heads]), scales[img], 0.5, 0.4)
        assert 0 < len(exp) <= k
        _check_faces(got[img], exp)


def test_k4_cap_truncates_in_score_order(ctx):
    rng = np.random.default_rng(12)
    heads = _random_heads(rng, 1, -3.0)
    full = ctx.scrfd_decode_nms(heads, np.array([1.0], np.float32), 0.5, 0.4, cap=4096)[0]
    few = ctx.scrfd_decode_nms(heads, np.array([1.0], np.float32), 0.5, 0.4, cap=5)[0]
    assert len(full) > 5 and len(few) == 5
    assert np.array_equal(few, full[:5])


def test_r4_l2_normalize_and_k8_compare(ctx, capi):
    rng = np.random.default_rng(2)
    x = rng.normal(size=(37, 512)).astype(np.float32)
    x[5] = 0
    y = ctx.l2_normalize(x)
    ref = orec.normalize_rows(x)
    assert np.allclose(y, ref, atol=1e-6) and np.array_equal(y[5], x[5])   # norm == 0 -> untouched
    for i in (0, 9):
        assert np.allclose(y[i], orec.normalize(x[i]), atol=1e-6)
    a, b = y[:16], y[16:32]
    sim = ctx.compare_batch(a, b)
    ref = np.array([orec.compare_faces(a[i], b[i]) for i in range(16)], np.float32)
    assert np.allclose(sim, ref, atol=1e-6)
    assert [orec.same_person(s) for s in sim] == [orec.same_person(s) for s in ref]
