"""Parity at BASELINE.json's full sizes: size-independent properties (determinism, independence of an
item's result from the batch around it, unit norms, planted matches, split invariance) PLUS the oracle
on a sample of the full-size batch (frames 0 / 63 of the 64, 8 of the 1024 crops, 16 of the 4096
queries against the rows pulled back from the 1.25 M-row shard) -- the oracle finishes those in seconds."""
import numpy as np
import pytest
import torch

import parity
from oracle import detector as odet
from oracle import gallery as ogal
from oracle import nets
from oracle import recognizer as orec

pytestmark = pytest.mark.gpu


def test_config2_scrfd_batch64_determinism_and_batch_invariance(ctx, det_wdict):
    """configs[1]: SCRFD det_500m, batch 64 synthetic 640x640 frames."""
    rng = np.random.default_rng(64)
    frames = [rng.integers(0, 256, (640, 640, 3), dtype=np.uint8) for _ in range(64)]
    a = ctx.detect_batch(frames, cap=1024)
    b = ctx.detect_batch(frames, cap=1024)
    assert all(np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(a, b))   # bit-identical reruns
    for i in (0, 17, 63):                      # a frame's detections do not depend on its batch
        one = ctx.detect(frames[i], cap=1024)
        assert np.array_equal(one.view(np.uint8), a[i].view(np.uint8))
    for d in a:                                # nms() leaves the list sorted by score (face_detector.cpp:357-383)
        assert np.all(np.diff(d["score"]) <= 0) and np.all(d["score"] > 0.5)
    # oracle on a sample of the batch, at a threshold where hundreds of candidates reach NMS
    low = ctx.detect_batch(frames, 0.02, 0.4, cap=4096)
    det = odet.FaceDetector(det_wdict)
    for i in (0, 63):
        chw, scale = odet.preprocess(frames[i])
        heads = det.run_network(chw)
        parity.assert_detections_explained(a[i], heads, scale, 0.5, 0.4)
        st = parity.assert_detections_explained(low[i], heads, scale, 0.02, 0.4)
        assert st["common"] > 20, st


def test_config3_arcface_batch1024_norms_and_batch_invariance(ctx, rec_wdict):
    """configs[2]: ArcFace w600k_r50, batch 1024 aligned 112x112 crops."""
    rng = np.random.default_rng(1024)
    crops = rng.integers(0, 256, (1024, 112, 112, 3), dtype=np.uint8)
    crops[777] = crops[5]                      # identical crops in different batch positions
    emb = ctx.embed_aligned(crops)
    assert emb.shape == (1024, 512) and np.isfinite(emb).all()
    assert np.abs(np.linalg.norm(emb, axis=1) - 1.0).max() < 1e-5
    assert np.array_equal(emb[777], emb[5])
    small = ctx.embed_aligned(crops[1000:1008])
    assert np.array_equal(small, emb[1000:1008])
    assert np.array_equal(ctx.embed_aligned(crops[:256]), emb[:256])
    # oracle on 8 random crops of the 1024
    pick = np.sort(rng.choice(1024, 8, replace=False))
    chw = np.stack([orec.preprocess(crops[i]) for i in pick])
    ref = orec.normalize_rows(nets.iresnet50_forward(rec_wdict, torch.from_numpy(chw)).numpy())
    cos = (emb[pick] * ref).sum(1)
    assert cos.min() >= 0.999, cos


def test_config5_gallery_full_shard_planted_and_split_invariance(ctx, capi):
    """configs[4] on one shard: 1.25 M x 512 rows, 4096 queries, top-10.  Planted rows come back
    as top-1 with score ~1, every list is sorted, and searching the shard as two half galleries
    + fr_topk_merge gives the same result as one search (what the multi-GPU merge relies on)."""
    n_rows, nq, k = 1_250_000, 4096, 10
    g = capi.Gallery(ctx, n_rows, index_base=0)
    g.fill_synthetic(n_rows, seed=1000)
    rng = np.random.default_rng(9)
    planted = np.sort(rng.choice(n_rows, 128, replace=False))
    q = rng.normal(size=(nq, 512)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    for j, r in enumerate(planted):
        q[j * 32] = g.get_rows(int(r), 1)[0]
    s, i = g.search(q, k)
    assert np.array_equal(i[::32, 0][:128], planted)
    assert np.all(s[::32, 0][:128] > 0.995)
    assert np.all(np.diff(s, axis=1) <= 0) and i.min() >= 0 and i.max() < n_rows
    # numpy oracle for 16 random queries against the whole shard, rows pulled back through the C ABI
    pick = np.sort(rng.choice(nq, 16, replace=False))
    best_s = np.full((16, k), -np.inf, np.float32)
    best_i = np.zeros((16, k), np.int64)
    chunk = 125_000
    for first in range(0, n_rows, chunk):
        rows = g.get_rows(first, chunk)
        cs, ci = ogal.topk(q[pick], rows, k)
        ms, mi = ogal.merge_topk([best_s, cs], [best_i, ci + first], k) if first else (cs, ci + first)
        best_s, best_i = ms, mi
    assert np.allclose(s[pick], best_s, atol=2e-5), float(np.abs(s[pick] - best_s).max())
    clear = np.ones_like(best_i, bool)
    dgap = np.abs(np.diff(best_s, axis=1)) > 1e-4
    clear[:, 1:] &= dgap
    clear[:, :-1] &= dgap
    assert np.array_equal(i[pick][clear], best_i[clear])
    # split invariance on a sub-range (two galleries holding rows [0, h) and [h, 2h))
    h = 150_000
    lo, hi = capi.Gallery(ctx, h, index_base=0), capi.Gallery(ctx, h, index_base=h)
    step = 50_000
    for first in range(0, h, step):
        lo.add(g.get_rows(first, step))
        hi.add(g.get_rows(h + first, step))
    both = capi.Gallery(ctx, 2 * h, index_base=0)
    for first in range(0, 2 * h, step):
        both.add(g.get_rows(first, step))
    s1, i1 = both.search(q[:512], k)
    sa, ia = lo.search(q[:512], k)
    sb, ib = hi.search(q[:512], k)
    sm, im = capi.topk_merge(ctx, np.stack([sa, sb]), np.stack([ia, ib]), k)
    assert np.array_equal(sm, s1) and np.array_equal(im, i1)
