"""End-to-end GPU parity: detect / extractFeature / fused pipeline / API mirror vs the oracle."""
import numpy as np
import pytest
import torch

from conftest import SEED, faces_from_landmarks, synth_landmarks
import parity
from oracle import detector as odet
from oracle import recognizer as orec

pytestmark = pytest.mark.gpu


def _frames(rng, n, h=640, w=640):
    return [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(n)]


def test_detect_end_to_end_vs_oracle(ctx, det_wdict):
    """fr_detect vs the oracle's detect (src/face_detector.cpp:139-222).  Every difference between the two
    lists must be EXPLAINED by a boundary event of the reference's arithmetic (tests/parity.py); and the
    list must be bit-identical to the oracle's decode + postprocess + NMS run on the GPU's own heads."""
    rng = np.random.default_rng(41)
    det = odet.FaceDetector(det_wdict)
    for thr in (0.5, 0.02):          # 0.02: hundreds of candidates per frame reach NMS with the seeded weights
        for im in _frames(rng, 2) + _frames(rng, 1, 480, 640):
            got = ctx.detect(im, thr, 0.4, cap=4096)
            chw, scale = odet.preprocess(im)
            heads = det.run_network(chw)
            st = parity.assert_detections_explained(got, heads, scale, thr, 0.4)
            assert st["common"] >= 1 or thr == 0.5
            # same frame through the stage hooks: K1 (bit-exact) -> K2 -> oracle decode/NMS on the GPU heads
            g_chw, g_scale = ctx.det_preprocess([im])
            assert np.array_equal(g_chw[0], chw) and g_scale[0] == scale
            g_heads = [h[0] for h in ctx.scrfd_forward(g_chw)]
            exp = odet.postprocess(odet.scrfd_decode(g_heads), scale, thr, 0.4)
            assert len(exp) == len(got)
            for r, e in zip(got, exp):
                assert (int(r["x"]), int(r["y"]), int(r["w"]), int(r["h"])) == (e.x, e.y, e.w, e.h)
                assert float(r["score"]) == float(e.score)
                assert np.array_equal(np.asarray(r["lm"], np.float32), e.landmarks.reshape(10))


def test_detect_batch_equals_single(ctx):
    rng = np.random.default_rng(42)
    ims = _frames(rng, 3) + _frames(rng, 1, 360, 640)
    batch = ctx.detect_batch(ims, 0.5, 0.4, cap=512)
    for im, b in zip(ims, batch):
        assert np.array_equal(ctx.detect(im, 0.5, 0.4, cap=512), b)


def test_extract_feature_end_to_end_vs_oracle(ctx, capi, rec_wdict):
    import cv2
    rng = np.random.default_rng(43)
    img = cv2.GaussianBlur(rng.integers(0, 256, (480, 640, 3), dtype=np.uint8), (0, 0), 1.2)
    lms = synth_landmarks(rng, 6, 640, 480)
    faces = faces_from_landmarks(capi, lms)
    emb, valid = ctx.embed_faces([img], faces, [0] * 6)
    assert valid.all()
    rec = orec.FaceRecognizer(rec_wdict)
    for i in range(6):
        fb = odet.FaceBox(int(faces[i]["x"]), int(faces[i]["y"]), int(faces[i]["w"]), int(faces[i]["h"]), 0.9, lms[i])
        ref = rec.extract_feature(img, fb)
        assert float((emb[i] * ref).sum()) >= 0.999


def test_extract_feature_simple_vs_oracle(ctx, rec_wdict):
    """fr_embed_simple vs FaceRecognizer::extractFeatureSimple (src/face_recognizer.cpp:152-234): the
    cv::resize(image -> 112x112) crop is bit-exact, the embedding within cosine 0.999, and the 0.6
    decision on pairs agrees (src/main.cpp:169-178)."""
    import cv2
    rng = np.random.default_rng(46)
    rec = orec.FaceRecognizer(rec_wdict)
    embs, refs = [], []
    for shape in ((480, 640, 3), (112, 112, 3), (37, 53, 3), (1000, 333, 3), (224, 224, 3)):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        if shape[0] > 200:
            img = cv2.GaussianBlur(img, (0, 0), 2.0)
        assert np.array_equal(ctx.resize_linear(img, 112, 112), cv2.resize(img, (112, 112)))
        e = ctx.embed_simple(img)
        r = rec.extract_feature_simple(img)
        assert abs(float(np.linalg.norm(e)) - 1.0) < 1e-5
        assert float((e * r).sum()) >= 0.999
        # the strided view a cv::Mat ROI would be (step != cols*3)
        big = np.zeros((shape[0] + 6, shape[1] + 10, 3), np.uint8)
        big[3:3 + shape[0], 5:5 + shape[1]] = img
        assert np.array_equal(ctx.embed_simple(big[3:3 + shape[0], 5:5 + shape[1]]), e)
        embs.append(e)
        refs.append(r)
    for i in range(len(embs)):
        for j in range(i + 1, len(embs)):
            s_ref = orec.compare_faces(refs[i], refs[j])
            if abs(float(s_ref) - 0.6) > 2e-3:
                assert orec.same_person(orec.compare_faces(embs[i], embs[j])) == orec.same_person(s_ref)


def test_pipeline_matches_staged_calls(ctx, capi):
    rng = np.random.default_rng(44)
    ims = _frames(rng, 4)
    K = 8
    pad = faces_from_landmarks(capi, synth_landmarks(rng, 4 * K, 640, 640)).reshape(4, K)
    faces, n_det, emb, valid = ctx.pipeline(ims, K, pad)
    dets = ctx.detect_batch(ims, 0.5, 0.4, cap=64)
    for i in range(4):
        assert n_det[i] == len(dets[i])
        sel = np.concatenate([dets[i][:K], pad[i][min(len(dets[i]), K):]])
        assert np.array_equal(faces[i], sel)
        e2, v2 = ctx.embed_faces([ims[i]], sel, [0] * K)
        assert np.array_equal(v2, valid[i])
        assert np.allclose(emb[i], e2, atol=1e-6)
    assert valid.all()


def test_api_mirror_compare_mode(capi, tmp_path):
    """compare-mode flow of src/main.cpp:67-123 through the reference-named classes."""
    from facerecognizeonnx_b200 import api
    api._Shared.seed = SEED
    det, rec = api.FaceDetector(), api.FaceRecognizer()
    assert det.detect(np.zeros((10, 10, 3), np.uint8)) == []          # "Model not loaded!"
    assert not det.loadModel(str(tmp_path / "det_500m.onnx"))           # absent -> false, like the reference
    assert not rec.loadModel(str(tmp_path / "w600k_r50.onnx"))          # (src/face_detector.cpp:86-89)
    assert det.detect(np.zeros((10, 10, 3), np.uint8)) == []            # still "Model not loaded!"
    api._Shared.allow_random_init = True                                # explicit opt-in (FR_ALLOW_RANDOM_INIT=1)
    assert det.loadModel(str(tmp_path / "det_500m.onnx"))
    assert rec.loadModel(str(tmp_path / "w600k_r50.onnx"))
    assert det.detect(np.zeros((0, 0, 3), np.uint8)) == []             # "Input image is empty!"
    rng = np.random.default_rng(45)
    im = rng.integers(0, 256, (640, 640, 3), dtype=np.uint8)
    faces = det.detect(im)
    assert all(isinstance(f, api.FaceBox) for f in faces)
    lm = synth_landmarks(rng, 2, 640, 640)
    f1 = api.FaceBox((10, 10, 100, 100), 0.9, lm[0])
    f2 = api.FaceBox((10, 10, 100, 100), 0.9, lm[1])
    e1, e2 = rec.extractFeature(im, f1), rec.extractFeature(im, f2)
    assert e1.shape == (512,) and abs(float(np.linalg.norm(e1)) - 1) < 1e-5
    s = rec.compareFaces(e1, e2)
    assert 0.0 <= s <= 1.0 and rec.compareFaces(e1, e1) > 0.999
    assert rec.compareFaces(e1, e2[:100]) == 0.0                       # size mismatch -> 0.0f
    bad = api.FaceBox((900, 900, 5, 5), 0.9, np.zeros((5, 2), np.float32))
    assert rec.extractFeature(im, bad).size == 0                       # alignment failed -> empty
    assert rec.extractFeatureSimple(im).shape == (512,)


def test_pipeline_submit_wait_matches_sync(ctx, capi):
    """fr_pipeline_submit / fr_pipeline_wait (two batches in flight, uploads on the copy stream)
    return exactly what the synchronous fr_pipeline_batch returns."""
    import torch
    rng = np.random.default_rng(77)
    K, n_img = 3, 4
    batches = [[rng.integers(0, 256, (480, 640, 3), dtype=np.uint8) for _ in range(n_img)] for _ in range(3)]
    lms = synth_landmarks(rng, n_img * K, 640, 480, outlier_frac=0.0)
    pad = faces_from_landmarks(capi, lms).reshape(n_img, K)
    ref = [ctx.pipeline(b, K, pad) for b in batches]

    def outs():
        return (np.zeros((n_img, K), capi.FACE_DTYPE), np.zeros(n_img, np.int32),
                np.zeros((n_img, K, 512), np.float32), np.zeros((n_img, K), np.int32))

    o = [outs() for _ in batches]
    t0 = ctx.pipeline_submit(batches[0], K, pad, *o[0])
    t1 = ctx.pipeline_submit(batches[1], K, pad, *o[1])
    ctx.pipeline_wait(t0)
    t2 = ctx.pipeline_submit(batches[2], K, pad, *o[2])
    ctx.pipeline_wait(t1)
    ctx.pipeline_wait(t2)
    for got, exp in zip(o, ref):
        assert np.array_equal(got[0].view(np.uint8), exp[0].view(np.uint8))
        assert np.array_equal(got[1], exp[1])
        assert np.array_equal(got[2], exp[2])
        assert np.array_equal(got[3], exp[3])


def test_detect_edge_cases_vs_oracle(ctx, capi, det_wdict):
    """The reference's guard rails and awkward inputs (src/face_detector.cpp:139-222): tiny and huge frames,
    extreme aspect ratios, ragged batches, no detection above the threshold, thresholds at their limits."""
    rng = np.random.default_rng(47)
    det = odet.FaceDetector(det_wdict)
    shapes = [(1, 1), (2, 3), (7, 640), (640, 7), (1080, 1920), (1920, 1080), (641, 639), (333, 517)]
    ims = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    CAP = 16800                                                 # every anchor: the cap must never truncate here (8225 candidates on the 1080p frame)
    batch = ctx.detect_batch(ims, 0.02, 0.4, cap=CAP)           # ragged batch: every image its own geometry
    for im, got in zip(ims, batch):
        chw, scale = odet.preprocess(im)
        assert chw is not None
        heads = det.run_network(chw)
        parity.assert_detections_explained(got, heads, scale, 0.02, 0.4)
        assert len(got) < CAP and np.array_equal(ctx.detect(im, 0.02, 0.4, cap=CAP).view(np.uint8), got.view(np.uint8))
    # a frame whose letterboxed size truncates to zero is refused like preprocess() does (:109-113) -> empty result
    flat = rng.integers(0, 256, (3, 2000, 3), dtype=np.uint8)
    assert odet.preprocess(flat)[0] is None
    with pytest.raises(capi.FrError) as e:
        ctx.detect(flat)
    assert e.value.code == capi.FR_ERR_INVALID_ARG
    # nothing above the threshold: zero detections, not an error; strict '>' at score_thr = 1.0
    assert len(ctx.detect(ims[4], 0.9999, 0.4)) == 0
    assert len(ctx.detect(ims[4], 1.0, 0.4)) == 0
    # nms_thr = 0: any overlap suppresses (strict '>' on an IoU of 0 keeps disjoint boxes); nms_thr >= 1: nothing is suppressed
    for nms_thr in (0.0, 1.0):
        got = ctx.detect(ims[4], 0.02, nms_thr, cap=CAP)
        chw, scale = odet.preprocess(ims[4])
        parity.assert_detections_explained(got, det.run_network(chw), scale, 0.02, nms_thr)
    # the fused pipeline with no detection and no padding faces marks every slot invalid and returns zeros
    faces, n_det, emb, valid = ctx.pipeline([ims[4]], 4, None, score_thr=0.9999)
    assert n_det[0] == 0 and not valid.any() and not emb.any()


def test_small_batch_graph_replay_is_transparent(ctx, capi):
    """fr_detect / fr_embed at a handful of images replay their launch chain as a CUDA graph from the third
    call with the same shapes on (capi.cu: run_graphed).  Eager run, capture, and replays must return the
    same bytes for the same input, different inputs of the same shape must not see each other's results,
    other shapes in between must not disturb a cached graph, and a reallocation of the scratch buffers (a
    larger batch) must retire the graphs instead of replaying stale pointers."""
    rng = np.random.default_rng(47)
    a, b = _frames(rng, 2)
    c = _frames(rng, 1, 480, 600)[0]
    l0 = ctx.launch_count()
    first_a = ctx.detect(a, 0.5, 0.4, cap=64)          # eager
    per_call = ctx.launch_count() - l0
    first_b = ctx.detect(b, 0.5, 0.4, cap=64)          # captured
    assert len(first_a) > 0 and not np.array_equal(first_a, first_b)
    for _ in range(3):                                 # replays
        l1 = ctx.launch_count()
        assert np.array_equal(ctx.detect(a, 0.5, 0.4, cap=64), first_a)
        assert ctx.launch_count() - l1 == per_call     # replays still count their kernels
        assert np.array_equal(ctx.detect(b, 0.5, 0.4, cap=64), first_b)
    first_c = ctx.detect(c, 0.5, 0.4, cap=64)          # another shape: its own key
    assert np.array_equal(ctx.detect(a, 0.5, 0.4, cap=64), first_a)
    assert np.array_equal(ctx.detect(c, 0.5, 0.4, cap=64), first_c)
    assert np.array_equal(ctx.detect(c, 0.5, 0.4, cap=64), first_c)
    assert np.array_equal(ctx.detect(a, 0.3, 0.4, cap=64), ctx.detect(a, 0.3, 0.4, cap=64))   # threshold is in the key
    assert np.array_equal(ctx.detect(a, 0.5, 0.4, cap=64), first_a)
    # embeddings of detected faces: eager / capture / replay
    fa, fb = first_a[:3].copy(), first_b[:2].copy()
    ea = [ctx.embed_faces([a], fa, [0] * len(fa))[0] for _ in range(4)]
    eb = [ctx.embed_faces([b], fb, [0] * len(fb))[0] for _ in range(3)]
    for e in ea[1:]:
        assert np.array_equal(e, ea[0])
    for e in eb[1:]:
        assert np.array_equal(e, eb[0])
    assert np.array_equal(ctx.embed_faces([a], fa, [0] * len(fa))[0], ea[0])
    # grow every scratch buffer (64 frames, 128 faces), then the small calls again
    big = _frames(rng, 6)
    ctx.detect_batch(big, 0.5, 0.4, cap=512)
    lms = synth_landmarks(rng, 200, 640, 640)
    ctx.embed_faces([a], faces_from_landmarks(capi, lms), [0] * 200)
    for _ in range(3):
        assert np.array_equal(ctx.detect(a, 0.5, 0.4, cap=64), first_a)
        assert np.array_equal(ctx.embed_faces([a], fa, [0] * len(fa))[0], ea[0])


def test_pipeline_graph_replay_is_transparent(ctx, capi):
    """fr_pipeline_batch on device-resident frames and fr_pipeline_submit replay the whole ~90-launch chain as a
    CUDA graph from the third call with the same buffers on.  The frame CONTENT changes between calls (same
    device buffers, new pixels): eager, captured and replayed calls must agree with each other for the same
    content, with the host-buffer entry point (which never uses a graph), and two submit slots must not mix."""
    rng = np.random.default_rng(48)
    n_img, K = 3, 4
    sets = [np.stack(_frames(rng, n_img)) for _ in range(2)]
    lms = synth_landmarks(rng, n_img * K, 640, 640)
    pad = faces_from_landmarks(capi, lms).reshape(n_img, K)
    ref = [ctx.pipeline([s[i] for i in range(n_img)], K, pad) for s in sets]       # host buffers: eager launches
    dev = torch.device("cuda", 0)
    frames_d = torch.empty((n_img, 640, 640, 3), dtype=torch.uint8, device=dev)
    pad_d = torch.from_numpy(pad.view(np.uint8).reshape(n_img * K, 60).copy()).to(dev)
    o_faces = torch.empty((n_img * K, 60), dtype=torch.uint8, device=dev)
    o_ndet = torch.empty(n_img, dtype=torch.int32, device=dev)
    o_emb = torch.empty((n_img * K, 512), dtype=torch.float32, device=dev)
    o_valid = torch.empty(n_img * K, dtype=torch.int32, device=dev)
    fb = 640 * 640 * 3
    for it in range(6):                                  # eager, capture, 4 replays; content alternates
        which = it % 2
        frames_d.copy_(torch.from_numpy(sets[which]))
        torch.cuda.synchronize()
        ctx.pipeline_dev([frames_d.data_ptr() + j * fb for j in range(n_img)], 640, 640, 640 * 3, K, pad_d.data_ptr(),
                         o_faces.data_ptr(), o_ndet.data_ptr(), o_emb.data_ptr(), o_valid.data_ptr())
        ctx.synchronize()
        faces, n_det, emb, valid = ref[which]
        assert np.array_equal(o_ndet.cpu().numpy(), n_det), it
        assert np.array_equal(o_emb.cpu().numpy().reshape(n_img, K, 512), emb), it
        assert np.array_equal(o_faces.cpu().numpy().view(capi.FACE_DTYPE).reshape(n_img, K), faces), it
        assert np.array_equal(o_valid.cpu().numpy().reshape(n_img, K), valid), it
    # streaming entry point: two slots in flight, content alternates, 8 batches (each slot: eager, capture, replays)
    outs = [(np.zeros((n_img, K), capi.FACE_DTYPE), np.zeros(n_img, np.int32), np.zeros((n_img, K, 512), np.float32),
             np.zeros((n_img, K), np.int32)) for _ in range(2)]
    tickets = []
    for it in range(8):
        which = (it // 2) % 2                            # AABBAABB: both slots see both contents
        o = outs[it % 2]
        if len(tickets) == 2:
            t_old, w_old, o_old = tickets.pop(0)
            ctx.pipeline_wait(t_old)
            assert np.array_equal(o_old[2], ref[w_old][2]) and np.array_equal(o_old[1], ref[w_old][1])
        tk = ctx.pipeline_submit([sets[which][i] for i in range(n_img)], K, pad, o[0], o[1], o[2], o[3])
        tickets.append((tk, which, o))
    for t_old, w_old, o_old in tickets:
        ctx.pipeline_wait(t_old)
        assert np.array_equal(o_old[2], ref[w_old][2]) and np.array_equal(o_old[0], ref[w_old][0])
