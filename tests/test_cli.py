"""The C++ drop-in: libface_api.so (FaceDetector / FaceRecognizer, reference src/face_detector.h:16-20,
src/face_recognizer.h:11-17) and the InsightFaceDemo CLI (reference src/main.cpp:264-319).

CPU: the binary exists, links, and fails the reference's way without a GPU (loadModel -> false,
message on stderr, exit code -1 = 255; src/main.cpp:274-283).
GPU: detect / compare / simple / webcam on synthetic PPM images print the reference's console
lines, and the numbers equal what the C ABI returns for the same images."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "facerecognizeonnx_b200")
CLI = os.path.join(PKG, "InsightFaceDemo")


def _build_host():
    subprocess.run(["make", "-C", os.path.join(PKG, "host")], check=True, capture_output=True)


def _run(args, cwd, random_init=True):
    """random_init: the tests have no model files, so they opt in to seeded random-init weights
    (FR_ALLOW_RANDOM_INIT=1); without the opt-in a missing model file fails like the reference."""
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = PKG + ":" + env.get("LD_LIBRARY_PATH", "")
    env.pop("FR_ALLOW_RANDOM_INIT", None)
    if random_init:
        env["FR_ALLOW_RANDOM_INIT"] = "1"
    return subprocess.run([CLI] + args, cwd=cwd, env=env, capture_output=True, text=True, timeout=300)


def _write_ppm(path, bgr):
    h, w, _ = bgr.shape
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (w, h))
        f.write(np.ascontiguousarray(bgr[:, :, ::-1]).tobytes())


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_cli_builds_and_fails_like_the_reference_without_gpu(tmp_path):
    _build_host()
    assert os.path.exists(CLI) and os.path.exists(os.path.join(PKG, "libface_api.so"))
    if _has_gpu():
        pytest.skip("GPU present: covered by the gpu tests")
    r = _run([], str(tmp_path))
    assert r.returncode == 255                      # main returns -1 when a model cannot be loaded
    assert "无法加载人脸检测模型" in r.stderr           # src/main.cpp:275
    assert "所有模型加载成功" not in r.stdout


def test_cli_missing_model_file_is_an_error_without_the_opt_in(tmp_path):
    """a missing / mistyped model path must fail like the reference (loadModel -> false,
    main -> -1, src/main.cpp:274-277), never fall back to random networks silently.  Holds with or
    without a GPU: the file check comes first."""
    _build_host()
    r = _run(["detect", "x.ppm"], str(tmp_path), random_init=False)
    assert r.returncode == 255
    assert "cannot open models/det_500m.onnx" in r.stderr and "无法加载人脸检测模型" in r.stderr
    assert "检测到" not in r.stdout


@pytest.mark.gpu
def test_cli_loads_real_onnx_files_from_models_dir(tmp_path, capi):
    """models/det_500m.onnx + models/w600k_r50.onnx present (raw, unfolded synthetic exports): the CLI
    loads them (no opt-in), and `compare` prints the similarity the C ABI gives with the same files."""
    import onnx_emit
    from oracle import weights as ow
    _build_host()
    os.makedirs(tmp_path / "models")
    pd, pr = str(tmp_path / "models" / "det_500m.onnx"), str(tmp_path / "models" / "w600k_r50.onnx")
    onnx_emit.emit_det_full(ow.trained_like_det(5), pd, raw=True, bbox_scales=(0.9, 1.7, 3.1))
    onnx_emit.emit_rec_full(ow.trained_like_rec(5), pr, raw=True)
    rng = np.random.default_rng(8)
    a = rng.integers(0, 256, (640, 640, 3), dtype=np.uint8)
    b = rng.integers(0, 256, (640, 640, 3), dtype=np.uint8)
    pa, pb = str(tmp_path / "a.ppm"), str(tmp_path / "b.ppm")
    _write_ppm(pa, a)
    _write_ppm(pb, b)
    r = _run(["compare", pa, pb], str(tmp_path), random_init=False)
    assert r.returncode == 0, r.stderr
    assert "所有模型加载成功" in r.stdout and "random-init" not in r.stderr
    c = capi.Context(0, capi.Weights(capi.FR_MODEL_DET, pd), capi.Weights(capi.FR_MODEL_REC, pr))
    fa, fb = c.detect(a, cap=1024), c.detect(b, cap=1024)
    assert len(fa) and len(fb)
    (ea,), _ = c.embed_faces([a], fa[:1], [0])
    (eb,), _ = c.embed_faces([b], fb[:1], [0])
    sim_cli = float(re.search(r"相似度: ([0-9.]+)", r.stdout).group(1))
    assert abs(sim_cli - capi.compare(ea, eb)) < 2e-6
    c.close()


@pytest.mark.gpu
def test_cli_modes_match_the_c_abi(tmp_path, ctx, capi):
    _build_host()
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    b = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    pa, pb = str(tmp_path / "a.ppm"), str(tmp_path / "b.ppm")
    _write_ppm(pa, a)
    _write_ppm(pb, b)

    r = _run([], str(tmp_path))                       # usage: exit 0 (src/main.cpp:300)
    assert r.returncode == 0 and "使用方法" in r.stdout
    r = _run(["bogus"], str(tmp_path))                # invalid mode: -1 (src/main.cpp:315)
    assert r.returncode == 255 and "无效的命令或参数" in r.stderr

    # detect: "检测到 N 个人脸" + one line per face (src/main.cpp:52-58); same faces as fr_detect
    r = _run(["detect", pa], str(tmp_path))
    assert r.returncode == 0, r.stderr
    n_cli = int(re.search(r"检测到 (\d+) 个人脸", r.stdout).group(1))
    faces = ctx.detect(a, cap=1024)
    assert n_cli == len(faces)
    m = re.search(r"位置\((-?\d+), (-?\d+), (-?\d+), (-?\d+)\)", r.stdout)
    if faces is not None and len(faces):
        f0 = faces[0]
        assert tuple(int(v) for v in m.groups()) == (int(f0["x"]), int(f0["y"]), int(f0["w"]), int(f0["h"]))

    # compare: first face of each image (src/main.cpp:88-123)
    r = _run(["compare", pa, pb], str(tmp_path))
    assert r.returncode == 0, r.stderr
    assert "特征维度: 512" in r.stdout
    sim_cli = float(re.search(r"相似度: ([0-9.]+)", r.stdout).group(1))
    fa, fb = ctx.detect(a, cap=1024), ctx.detect(b, cap=1024)
    (ea,), _ = ctx.embed_faces([a], fa[:1], [0])
    (eb,), _ = ctx.embed_faces([b], fb[:1], [0])
    sim = capi.compare(ea, eb)
    assert abs(sim_cli - sim) < 2e-6
    assert ("同一人" in r.stdout) == (sim > 0.6)

    # simple: no detector, resize whole image (src/main.cpp:136-199)
    r = _run(["simple", pa, pb], str(tmp_path))
    assert r.returncode == 0, r.stderr
    sim_cli = float(re.search(r"相似度: ([0-9.]+)", r.stdout).group(1))
    sim = capi.compare(ctx.embed_simple(a), ctx.embed_simple(b))
    assert abs(sim_cli - sim) < 2e-6

    # missing file: message, no crash (src/main.cpp:43-46)
    r = _run(["detect", str(tmp_path / "nope.ppm")], str(tmp_path))
    assert r.returncode == 0 and "无法读取图像" in (r.stdout + r.stderr)


@pytest.mark.gpu
def test_cli_webcam_frame_list(tmp_path, ctx, capi):
    """webcam mode over a frame list (src/main.cpp:201-262): the first frame with a face sets the
    reference, every face of the later frames is matched against it with the 0.6 rule; the
    similarities equal fr_compare on the C-ABI embeddings (batched extractFeatures path)."""
    _build_host()
    rng = np.random.default_rng(4)
    frames = [rng.integers(0, 256, (480, 640, 3), dtype=np.uint8) for _ in range(2)]
    paths = []
    for i, f in enumerate(frames):
        p = str(tmp_path / f"f{i}.ppm")
        _write_ppm(p, f)
        paths.append(p)
    r = _run(["webcam"] + paths, str(tmp_path))
    assert r.returncode == 0, r.stderr
    assert "已保存参考人脸特征" in r.stdout and "Reference set" in r.stdout
    sims = [float(m) for m in re.findall(r"Sim: ([0-9.eE+-]+)", r.stdout)]
    f0, f1 = ctx.detect(frames[0], cap=1024), ctx.detect(frames[1], cap=1024)
    assert len(sims) == len(f1)
    (ref,), _ = ctx.embed_faces([frames[0]], f0[:1], [0])
    emb, valid = ctx.embed_faces([frames[1]], f1, [0] * len(f1))
    exp = [capi.compare(ref, e) for e in emb]
    assert np.allclose(sims, exp, atol=2e-5)
    labels = re.findall(r"face \d+: (Match|Unknown)", r.stdout)
    assert labels == ["Match" if s > 0.6 else "Unknown" for s in exp]
