"""Integer / fixed-point restatements of the OpenCV calls on the hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  These are the recipes the CUDA
kernels implement; ``tests/test_oracle_cv.py`` pins each of them bit-for-bit
(or to 1e-6 for the double-precision estimator) against the ``cv2`` 4.13 wheel,
which is the same OpenCV code the reference links against
(``cv::resize`` src/face_detector.cpp:117, src/face_recognizer.cpp:123,170;
``cv::warpAffine`` src/face_recognizer.cpp:130;
``cv::estimateAffinePartial2D`` src/face_recognizer.cpp:110-113).
"""
from __future__ import annotations

import math

import numpy as np

# --------------------------------------------------------------------------
# cv::resize(src, dst, Size(nw, nh)), INTER_LINEAR, CV_8UC3
# --------------------------------------------------------------------------

INTER_RESIZE_COEF_BITS = 11
INTER_RESIZE_COEF_SCALE = 1 << INTER_RESIZE_COEF_BITS  # 2048


def _resize_axis_tables(n_dst: int, n_src: int, zero_frac_at_border: bool):
    """Per-destination-index source index pair and 11-bit weights.

    Horizontal tables zero the fraction at the borders; vertical tables keep
    the fraction and clip the indices (OpenCV's asymmetry)."""
    scale = 1.0 / (float(n_dst) / float(n_src))  # double, as cv::resize computes inv_scale
    i0 = np.empty(n_dst, np.int32)
    i1 = np.empty(n_dst, np.int32)
    w0 = np.empty(n_dst, np.int32)
    w1 = np.empty(n_dst, np.int32)
    for d in range(n_dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(math.floor(float(f)))
        f = np.float32(f - np.float32(s))
        if zero_frac_at_border:
            if s < 0:
                f = np.float32(0.0)
                s = 0
            if s >= n_src - 1:
                f = np.float32(0.0)
                s = n_src - 1
            s1 = min(s + 1, n_src - 1)
            s0 = s
        else:
            s0 = min(max(s, 0), n_src - 1)
            s1 = min(max(s + 1, 0), n_src - 1)
        # cvRound == round-half-to-even on the float product; saturate_cast<short>
        a0 = int(np.rint(np.float32(np.float32(1.0) - f) * np.float32(INTER_RESIZE_COEF_SCALE)))
        a1 = int(np.rint(f * np.float32(INTER_RESIZE_COEF_SCALE)))
        i0[d], i1[d], w0[d], w1[d] = s0, s1, a0, a1
    return i0, i1, w0, w1


def resize_linear_u8(src: np.ndarray, nw: int, nh: int) -> np.ndarray:
    """Bit-exact cv::resize(INTER_LINEAR) for uint8 HxWxC images."""
    assert src.dtype == np.uint8 and src.ndim == 3
    sh, sw = src.shape[:2]
    if (sw, sh) == (nw, nh):
        return src.copy()
    x0, x1, a0, a1 = _resize_axis_tables(nw, sw, True)
    y0, y1, b0, b1 = _resize_axis_tables(nh, sh, False)
    s = src.astype(np.int32)
    # horizontal pass: int32, scale 2^11
    hbuf = s[:, x0, :] * a0[None, :, None] + s[:, x1, :] * a1[None, :, None]
    r0 = hbuf[y0] >> 4
    r1 = hbuf[y1] >> 4
    out = (((b0[:, None, None] * r0) >> 16) + ((b1[:, None, None] * r1) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


# --------------------------------------------------------------------------
# cv::warpAffine(src, dst, M, Size(w,h)), INTER_LINEAR, BORDER_CONSTANT(0)
# --------------------------------------------------------------------------

AB_BITS = 10
AB_SCALE = 1 << AB_BITS
INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS


def invert_affine(M: np.ndarray) -> np.ndarray:
    """cv::warpAffine's internal inversion of the 2x3 double matrix."""
    M = np.asarray(M, np.float64)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11 = M[1, 1] * D
    A22 = M[0, 0] * D
    A12 = -M[0, 1] * D
    A21 = -M[1, 0] * D
    b1 = -A11 * M[0, 2] - A12 * M[1, 2]
    b2 = -A21 * M[0, 2] - A22 * M[1, 2]
    return np.array([[A11, A12, b1], [A21, A22, b2]], np.float64)


def _sat_i32(v: np.ndarray) -> np.ndarray:
    return np.clip(v, -2147483648.0, 2147483647.0).astype(np.int64)


def warp_affine_u8(src: np.ndarray, M: np.ndarray, dsize=(112, 112)) -> np.ndarray:
    """Bit-exact cv::warpAffine (fixed-point bilinear, constant 0 border)."""
    assert src.dtype == np.uint8 and src.ndim == 3
    dw, dh = dsize
    sh, sw, ch = src.shape
    inv = invert_affine(M)
    xs = np.arange(dw, dtype=np.float64)
    ys = np.arange(dh, dtype=np.float64)
    adelta = _sat_i32(np.rint(inv[0, 0] * xs * AB_SCALE))
    bdelta = _sat_i32(np.rint(inv[1, 0] * xs * AB_SCALE))
    rd = AB_SCALE // INTER_TAB_SIZE // 2  # 16
    X0 = _sat_i32(np.rint((inv[0, 1] * ys + inv[0, 2]) * AB_SCALE)) + rd
    Y0 = _sat_i32(np.rint((inv[1, 1] * ys + inv[1, 2]) * AB_SCALE)) + rd
    # int32 wrap-around addition, arithmetic shift
    X = ((X0[:, None] + adelta[None, :]).astype(np.int32)) >> (AB_BITS - INTER_BITS)
    Y = ((Y0[:, None] + bdelta[None, :]).astype(np.int32)) >> (AB_BITS - INTER_BITS)
    ix = np.clip(X >> INTER_BITS, -32768, 32767).astype(np.int64)
    iy = np.clip(Y >> INTER_BITS, -32768, 32767).astype(np.int64)
    fx = (X & (INTER_TAB_SIZE - 1)).astype(np.int64)
    fy = (Y & (INTER_TAB_SIZE - 1)).astype(np.int64)
    w00 = (32 - fx) * (32 - fy) * 32
    w10 = fx * (32 - fy) * 32
    w01 = (32 - fx) * fy * 32
    w11 = fx * fy * 32

    def tap(yy, xx):
        ok = (xx >= 0) & (xx < sw) & (yy >= 0) & (yy < sh)
        v = src[np.clip(yy, 0, sh - 1), np.clip(xx, 0, sw - 1)].astype(np.int64)
        return v * ok[..., None]

    acc = (w00[..., None] * tap(iy, ix) + w10[..., None] * tap(iy, ix + 1)
           + w01[..., None] * tap(iy + 1, ix) + w11[..., None] * tap(iy + 1, ix + 1))
    out = (acc + (1 << 14)) >> 15
    return out.astype(np.uint8)


# --------------------------------------------------------------------------
# cv::estimateAffinePartial2D(from, to) with all defaults
#   RANSAC (2-point model, thr 3.0, conf 0.99, <=2000 iters) -> inliers ->
#   Levenberg-Marquardt refine (== closed-form least squares at convergence)
# --------------------------------------------------------------------------

_RNG_MULT = 4164903690
_MASK64 = (1 << 64) - 1


class CvRNG:
    """cv::RNG (multiply-with-carry), as used by RANSACPointSetRegistrator."""

    def __init__(self, state: int = _MASK64):
        self.state = state & _MASK64

    def next(self) -> int:
        self.state = ((self.state & 0xFFFFFFFF) * _RNG_MULT + (self.state >> 32)) & _MASK64
        return self.state & 0xFFFFFFFF

    def uniform(self, a: int, b: int) -> int:
        return self.next() % (b - a) + a


def ransac_pair_table(n: int = 64, count: int = 5):
    """The data-independent sequence of 2-point samples RANSAC draws."""
    rng = CvRNG()
    out = []
    for _ in range(n):
        i0 = rng.uniform(0, count)
        while True:
            i1 = rng.uniform(0, count)
            if i1 != i0:
                break
        out.append((i0, i1))
    return out


def _ransac_update_niters(p: float, ep: float, model_points: int, max_iters: int) -> int:
    p = max(p, 0.0)
    p = min(p, 1.0)
    ep = max(ep, 0.0)
    ep = min(ep, 1.0)
    num = max(1.0 - p, np.finfo(np.float64).tiny)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < np.finfo(np.float64).tiny:
        return 0
    num = math.log(num)
    denom = math.log(denom)
    if denom >= 0 or -num >= max_iters * (-denom):
        return max_iters
    return int(np.rint(num / denom))


def _two_point_similarity(f0, f1, t0, t1):
    """AffinePartial2DEstimatorCallback::runKernel (double)."""
    x1, y1 = float(f0[0]), float(f0[1])
    x2, y2 = float(f1[0]), float(f1[1])
    X1, Y1 = float(t0[0]), float(t0[1])
    X2, Y2 = float(t1[0]), float(t1[1])
    with np.errstate(all="ignore"):
        d = np.float64(1.0) / np.float64((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2))
        S0 = d * ((X1 - X2) * (x1 - x2) + (Y1 - Y2) * (y1 - y2))
        S1 = d * ((Y1 - Y2) * (x1 - x2) - (X1 - X2) * (y1 - y2))
        S2 = d * ((Y1 - Y2) * (x1 * y2 - x2 * y1) - (X1 * y2 - X2 * y1) * (y1 - y2)
                  - (X1 * x2 - X2 * x1) * (x1 - x2))
        S3 = d * (-(X1 - X2) * (x1 * y2 - x2 * y1) - (Y1 * x2 - Y2 * x1) * (x1 - x2)
                  - (Y1 * y2 - Y2 * y1) * (y1 - y2))
    return S0, S1, S2, S3


def ransac_inliers(src: np.ndarray, dst: np.ndarray, thr: float = 3.0, conf: float = 0.99,
                   max_iters: int = 2000):
    """Replica of RANSACPointSetRegistrator::run for the 2-pt similarity model.

    Returns (mask[bool n], best_model(a,b,tx,ty) or None)."""
    src = np.asarray(src, np.float32).reshape(-1, 2)
    dst = np.asarray(dst, np.float32).reshape(-1, 2)
    n = src.shape[0]
    rng = CvRNG()
    niters = max_iters
    max_good = 0
    best_mask = np.zeros(n, bool)
    best_model = None
    thr2 = np.float32(thr * thr)
    it = 0
    while it < niters:
        i0 = rng.uniform(0, n)
        while True:
            i1 = rng.uniform(0, n)
            if i1 != i0:
                break
        a, b, tx, ty = _two_point_similarity(src[i0], src[i1], dst[i0], dst[i1])
        # computeError: model is double 2x3 but coefficients are cast to float
        with np.errstate(all="ignore"):
            F0, F1, F2 = np.float32(a), np.float32(-b), np.float32(tx)
            F3, F4, F5 = np.float32(b), np.float32(a), np.float32(ty)
            ex = F0 * src[:, 0] + F1 * src[:, 1] + F2 - dst[:, 0]
            ey = F3 * src[:, 0] + F4 * src[:, 1] + F5 - dst[:, 1]
            err = ex * ex + ey * ey
            mask = err <= thr2
        good = int(mask.sum())
        if good > max(max_good, 1):
            best_mask = mask.copy()
            best_model = (a, b, tx, ty)
            max_good = good
            niters = _ransac_update_niters(conf, float(n - good) / n, 2, niters)
        it += 1
    if max_good == 0:
        return best_mask, None
    return best_mask, best_model


def similarity_lstsq(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """Closed-form least-squares similarity [[a,-b,tx],[b,a,ty]] (double).

    This is the fixed point OpenCV's 10-iteration LM refine converges to."""
    src = np.asarray(src, np.float64).reshape(-1, 2)
    dst = np.asarray(dst, np.float64).reshape(-1, 2)
    n = src.shape[0]
    mx, my = src[:, 0].sum() / n, src[:, 1].sum() / n
    mX, mY = dst[:, 0].sum() / n, dst[:, 1].sum() / n
    sx, sy = src[:, 0] - mx, src[:, 1] - my
    sX, sY = dst[:, 0] - mX, dst[:, 1] - mY
    den = (sx * sx + sy * sy).sum()
    a = (sx * sX + sy * sY).sum() / den
    b = (sx * sY - sy * sX).sum() / den
    tx = mX - (a * mx - b * my)
    ty = mY - (b * mx + a * my)
    return np.array([[a, -b, tx], [b, a, ty]], np.float64)


def estimate_affine_partial_2d(src: np.ndarray, dst: np.ndarray):
    """Replica of cv::estimateAffinePartial2D(src, dst) with default arguments.

    Returns (M 2x3 float64 or None, inlier mask)."""
    mask, model = ransac_inliers(src, dst)
    if model is None:
        return None, mask
    src = np.asarray(src, np.float32).reshape(-1, 2)
    dst = np.asarray(dst, np.float32).reshape(-1, 2)
    if int(mask.sum()) <= 2:
        a, b, tx, ty = model
        if int(mask.sum()) == 2:
            return similarity_lstsq(src[mask], dst[mask]), mask
        return np.array([[a, -b, tx], [b, a, ty]], np.float64), mask
    return similarity_lstsq(src[mask], dst[mask]), mask
