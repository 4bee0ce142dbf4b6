"""Pins oracle/cv_recipes.py (the recipes the CUDA kernels implement) against cv2 itself --
the same OpenCV code the reference calls (src/face_detector.cpp:117, src/face_recognizer.cpp:
110-113,130).  CPU only."""
import cv2
import numpy as np
import pytest

from oracle import cv_recipes as R
from oracle import recognizer as orec

RESIZE_CASES = [(640, 480, 640, 480), (1280, 720, 640, 360), (1920, 1080, 640, 360), (517, 333, 640, 412),
                (800, 1000, 512, 640), (200, 300, 426, 640), (48, 64, 480, 640), (300, 200, 112, 112),
                (37, 91, 112, 112), (1, 1, 112, 112), (640, 640, 640, 640)]


@pytest.mark.parametrize("sw,sh,nw,nh", RESIZE_CASES)
def test_resize_bit_exact(sw, sh, nw, nh):
    rng = np.random.default_rng(sw * 7 + sh)
    img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
    assert np.array_equal(cv2.resize(img, (nw, nh)), R.resize_linear_u8(img, nw, nh))


def test_warp_affine_bit_exact():
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    for _ in range(40):
        s, th = rng.uniform(0.3, 1.5), rng.uniform(-0.5, 0.5)
        M = np.array([[s * np.cos(th), -s * np.sin(th), rng.uniform(-150, 150)],
                      [s * np.sin(th), s * np.cos(th), rng.uniform(-150, 150)]])
        assert np.array_equal(cv2.warpAffine(img, M, (112, 112)), R.warp_affine_u8(img, M, (112, 112)))


def test_warp_affine_far_outside_is_zero():
    img = np.full((50, 60, 3), 200, np.uint8)
    M = np.array([[1.0, 0.0, 5000.0], [0.0, 1.0, 5000.0]])
    assert np.array_equal(cv2.warpAffine(img, M, (112, 112)), R.warp_affine_u8(img, M, (112, 112)))


def test_ransac_pair_table_is_data_independent():
    assert R.ransac_pair_table(8) == [(0, 4), (0, 3), (1, 2), (1, 0), (3, 0), (0, 4), (1, 4), (3, 1)]


def test_estimate_affine_partial_matches_cv2():
    from conftest import synth_landmarks
    rng = np.random.default_rng(5)
    lms = synth_landmarks(rng, 600, outlier_frac=0.5)
    worst = 0.0
    for lm in lms:
        Mc, inl = cv2.estimateAffinePartial2D(lm, orec.TEMPLATE)
        Mo, mask = R.estimate_affine_partial_2d(lm, orec.TEMPLATE)
        assert (Mc is None) == (Mo is None)
        if Mc is None:
            continue
        assert np.array_equal(inl.ravel().astype(bool), mask)
        worst = max(worst, float(np.abs(Mc - Mo).max()))
    assert worst < 1e-7


def test_estimate_degenerate_returns_none():
    z = np.zeros((5, 2), np.float32)
    assert cv2.estimateAffinePartial2D(z, orec.TEMPLATE)[0] is None
    assert R.estimate_affine_partial_2d(z, orec.TEMPLATE)[0] is None
