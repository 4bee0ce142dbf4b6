import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 1  # weight seed shared by every test and by bench.py


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def capi():
    from facerecognizeonnx_b200 import capi as m
    m.lib()  # raises ImportError loudly if the .so is not built
    return m


@pytest.fixture(scope="session")
def det_weights(capi):
    return capi.Weights(capi.FR_MODEL_DET, None, SEED)


@pytest.fixture(scope="session")
def rec_weights(capi):
    return capi.Weights(capi.FR_MODEL_REC, None, SEED)


@pytest.fixture(scope="session")
def det_wdict(det_weights):
    return det_weights.to_dict()


@pytest.fixture(scope="session")
def rec_wdict(rec_weights):
    return rec_weights.to_dict()


@pytest.fixture(scope="session")
def ctx(capi, det_weights, rec_weights):
    c = capi.Context(0, det_weights, rec_weights)
    yield c
    c.close()


def synth_landmarks(rng, n, img_w=640, img_h=480, outlier_frac=0.3):
    """template x random similarity (+ jitter, + occasional outliers): SURVEY 8d generator."""
    tmpl = np.array([[38.2946, 51.6963], [73.5318, 51.5014], [56.0252, 71.7366],
                     [41.5493, 92.3655], [70.7299, 92.2041]], np.float32)
    out = np.zeros((n, 5, 2), np.float32)
    for i in range(n):
        s = rng.uniform(0.5, 4.0)
        th = rng.uniform(-0.6, 0.6)
        A = np.array([[s * np.cos(th), -s * np.sin(th)], [s * np.sin(th), s * np.cos(th)]])
        t = np.array([rng.uniform(0, img_w * 0.6), rng.uniform(0, img_h * 0.6)])
        pts = tmpl @ A.T + t + rng.normal(0, 0.5 * s, (5, 2))
        if rng.uniform() < outlier_frac:
            pts[rng.integers(0, 5)] += rng.normal(0, 20 * s, 2)
        out[i] = pts.astype(np.float32)
    return out


def faces_from_landmarks(capi, lms):
    n = lms.shape[0]
    f = np.zeros(n, capi.FACE_DTYPE)
    for i in range(n):
        x0, y0 = lms[i].min(0)
        x1, y1 = lms[i].max(0)
        f[i]["x"], f[i]["y"] = int(x0), int(y0)
        f[i]["w"], f[i]["h"] = int(x1 - x0) + 1, int(y1 - y0) + 1
        f[i]["score"] = 0.9
        f[i]["lm"] = lms[i].reshape(10)
    return f
