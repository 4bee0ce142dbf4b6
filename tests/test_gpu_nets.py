"""GPU parity tests for the two networks through the C ABI.

K6/K7 (IResNet-50, bf16 tensor cores, fp32 accumulate): cosine >= 0.999 vs the torch fp32
oracle on shared weights (north_star tolerance), per-block taps within bf16 error.
K2 (SCRFD, tcgen05 tf32 three-term split = fp32-grade products): bbox / kps heads within
1e-3 px after decode (error in stride units x stride), every intermediate activation within
2e-5 of its range; the decode stage itself is bit-exact given identical
heads (tests/test_gpu_stages.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import nets

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-12))


def _bf16(x):
    return torch.from_numpy(np.ascontiguousarray(x, np.float32)).bfloat16().float()


@pytest.mark.parametrize("n,cin,cout,h,w", [(2, 64, 64, 8, 8), (1, 64, 64, 15, 13), (2, 128, 128, 14, 14),
                                            (1, 64, 128, 9, 9), (1, 256, 256, 7, 7), (1, 128, 512, 7, 7),
                                            # 2-D tile mode (7x16 patches at 112^2, 8x14 at 56^2)
                                            (2, 64, 64, 56, 56), (1, 64, 64, 112, 112), (1, 64, 128, 56, 56)])
def test_tc_conv3x3_plain(ctx, n, cin, cout, h, w):
    rng = np.random.default_rng(cin + cout + h)
    x = rng.normal(size=(n, cin, h, w)).astype(np.float32)
    wt = (rng.normal(size=(cout, cin, 3, 3)) / np.sqrt(9 * cin)).astype(np.float32)
    b = rng.normal(size=cout).astype(np.float32)
    y = ctx.test_conv(x, wt, bias=b)
    ref = F.conv2d(_bf16(x), _bf16(wt), torch.from_numpy(b), padding=1).numpy()
    assert _rel(y, ref) < 6e-3, _rel(y, ref)   # output itself is rounded to bf16 (2^-9 relative)


@pytest.mark.parametrize("n,cin,cout,h,w", [(1, 256, 256, 14, 14), (3, 256, 256, 14, 14), (40, 256, 256, 14, 14),
                                            (5, 512, 512, 7, 7), (2, 128, 256, 7, 7), (90, 256, 256, 7, 7)])
def test_tc_conv3x3_two_cta_pair(ctx, n, cin, cout, h, w):
    """halo_gemm2_kernel (tcgen05.mma.cta_group::2, M = 256 over a CTA pair, each CTA loading half of the
    weight tile): all 256 / 512-output-channel 3x3 stride-1 layers.  Odd numbers of 128-row tiles (the peer
    CTA's half of the last item is past the end), more items than clusters (ring wrap-around, both TMEM
    buffers), and the N-split tail all appear in these shapes.  Bias + PReLU + residual epilogue."""
    rng = np.random.default_rng(n + cin + h)
    x = rng.normal(size=(n, cin, h, w)).astype(np.float32)
    wt = (rng.normal(size=(cout, cin, 3, 3)) / np.sqrt(9 * cin)).astype(np.float32)
    b = rng.normal(size=cout).astype(np.float32)
    slope = rng.uniform(0.1, 0.4, cout).astype(np.float32)
    res = rng.normal(size=(n, cout, h, w)).astype(np.float32)
    y = ctx.test_conv(x, wt, bias=b, prelu=slope, residual=res)
    ref = F.conv2d(_bf16(x), _bf16(wt), torch.from_numpy(b), padding=1)
    ref = (F.prelu(ref, torch.from_numpy(slope)) + _bf16(res)).numpy()
    assert _rel(y, ref) < 8e-3, _rel(y, ref)
    # column halves must not be swapped between the two CTAs: check the two N halves separately
    hn = cout // 2
    assert _rel(y[:, :hn], ref[:, :hn]) < 8e-3 and _rel(y[:, hn:], ref[:, hn:]) < 8e-3


def test_tc_conv3x3_full_epilogue(ctx):
    rng = np.random.default_rng(7)
    n, cin, cout, h, w = 2, 128, 128, 14, 14
    x = rng.normal(size=(n, cin, h, w)).astype(np.float32)
    wt = (rng.normal(size=(cout, cin, 3, 3)) / np.sqrt(9 * cin)).astype(np.float32)
    b = rng.normal(size=cout).astype(np.float32)
    sc = rng.uniform(0.5, 1.5, cin).astype(np.float32)
    sh = rng.normal(size=cin).astype(np.float32)
    slope = rng.uniform(0.1, 0.4, cout).astype(np.float32)
    res = rng.normal(size=(n, cout, h, w)).astype(np.float32)
    y = ctx.test_conv(x, wt, pre_scale=sc, pre_shift=sh, bias=b, prelu=slope, residual=res)
    xin = _bf16(x) * torch.from_numpy(sc)[None, :, None, None] + torch.from_numpy(sh)[None, :, None, None]
    ref = F.conv2d(xin, torch.from_numpy(wt), torch.from_numpy(b), padding=1)
    ref = F.prelu(ref, torch.from_numpy(slope)) + _bf16(res)
    assert _rel(y, ref.numpy()) < 8e-3, _rel(y, ref.numpy())


def test_tc_conv1x1(ctx):
    rng = np.random.default_rng(8)
    x = rng.normal(size=(2, 128, 9, 9)).astype(np.float32)
    wt = (rng.normal(size=(64, 128, 1, 1)) / np.sqrt(128)).astype(np.float32)
    y = ctx.test_conv(x, wt)
    ref = F.conv2d(_bf16(x), _bf16(wt)).numpy()
    assert _rel(y, ref) < 6e-3


def test_k6_iresnet_taps_and_output(ctx, rec_wdict):
    rng = np.random.default_rng(21)
    n = 3
    x = ((rng.integers(0, 256, (n, 3, 112, 112)).astype(np.float32)) - 127.5) / 128
    got = ctx.iresnet_forward(x)
    ref, taps = nets.iresnet50_forward(rec_wdict, torch.from_numpy(x), return_taps=True)
    ref = ref.numpy()
    names = ["stem"] + [f"l{li}.{bi}" for li, (nb, _) in enumerate(nets.REC_LAYERS) for bi in range(nb)]
    worst = []
    for bi, name in enumerate(names):
        t = taps[name].numpy()
        g = ctx.iresnet_tap(0 if bi == 0 else 2 * bi, n, t.shape[1:])
        worst.append((name, _rel(g, t)))
        if bi > 0:
            th = taps[name + ".h"].numpy()
            gh = ctx.iresnet_tap(2 * bi - 1, n, th.shape[1:])
            worst.append((name + ".h", _rel(gh, th)))
    bad = [w for w in worst if not w[1] < 3e-2]
    assert not bad, bad[:5]
    cos = (got * ref).sum(1) / (np.linalg.norm(got, axis=1) * np.linalg.norm(ref, axis=1))
    assert cos.min() >= 0.999, (cos, worst[-3:])


def test_k6_embed_aligned_batch_cosine_and_batch_invariance(ctx, rec_wdict):
    from oracle import recognizer as orec
    rng = np.random.default_rng(22)
    crops = rng.integers(0, 256, (9, 112, 112, 3), dtype=np.uint8)
    emb = ctx.embed_aligned(crops)
    assert np.allclose(np.linalg.norm(emb, axis=1), 1.0, atol=1e-5)
    chw = np.stack([orec.preprocess(c) for c in crops])
    ref = orec.normalize_rows(nets.iresnet50_forward(rec_wdict, torch.from_numpy(chw)).numpy())
    cos = (emb * ref).sum(1)
    assert cos.min() >= 0.999, cos
    one = ctx.embed_aligned(crops[4:5])
    assert np.array_equal(one[0], emb[4])   # per-item result independent of the batch around it
    # same-person decisions at the 0.6 threshold agree with the oracle for every pair
    for i in range(9):
        for j in range(i + 1, 9):
            s_ref = orec.compare_faces(ref[i], ref[j])
            if abs(float(s_ref) - 0.6) > 2e-3:
                assert orec.same_person(orec.compare_faces(emb[i], emb[j])) == orec.same_person(s_ref)


def test_k6_stem_mma_matches_oracle_and_simt_stem(ctx, rec_wdict):
    """The u8 hot path runs the stem on warp-level MMA (bf16 hi+lo weights, exact bf16 inputs); the fp32
    CHW entry point keeps the CUDA-core stem.  Both must give the oracle's stem activation up to the
    bf16 rounding of the output, and each other's up to one bf16 ulp plus the 16-bit weight split's
    absolute error (27 products x 2^-17)."""
    from oracle import recognizer as orec
    rng = np.random.default_rng(23)
    n = 5
    crops = rng.integers(0, 256, (n, 112, 112, 3), dtype=np.uint8)
    crops[0] = 0
    crops[1] = 255
    chw = np.stack([orec.preprocess(c) for c in crops])
    _, taps = nets.iresnet50_forward(rec_wdict, torch.from_numpy(chw), return_taps=True)
    ref = taps["stem"].numpy()
    ctx.embed_aligned(crops)
    got_mma = ctx.iresnet_tap(0, n, ref.shape[1:])
    ctx.iresnet_forward(chw)
    got_simt = ctx.iresnet_tap(0, n, ref.shape[1:])
    assert _rel(got_mma, ref) < 6e-3 and _rel(got_simt, ref) < 6e-3
    ulp = np.abs(got_simt) * 2.0 ** -7 + 2e-5
    assert (np.abs(got_mma - got_simt) <= ulp).all()
    assert (got_mma != got_simt).mean() < 0.02


def test_k6_embeddings_do_not_depend_on_batch_size_or_tail_split(ctx):
    """The conv kernel splits the last partial wave of tiles along N (tc::Params::tail_split); which tiles
    are split depends on the batch size.  A face's embedding must be bit-identical whatever batch it is in
    (each output element accumulates over K in the same order in a whole tile and in an N slice)."""
    rng = np.random.default_rng(24)
    crops = rng.integers(0, 256, (300, 112, 112, 3), dtype=np.uint8)
    full = ctx.embed_aligned(crops)
    for lo, hi in ((0, 37), (37, 187), (187, 300), (5, 6)):
        part = ctx.embed_aligned(crops[lo:hi])
        assert np.array_equal(part, full[lo:hi]), (lo, hi)


def test_k2_scrfd_heads_vs_oracle(ctx, det_wdict):
    """tcgen05 tf32 three-term split vs torch fp32: every activation tap and all 9 heads.  The
    bar is north_star's: boxes and landmarks within 1e-3 px after decode, i.e. the bbox / kps
    heads (stride units) within 1e-3 / stride; scores within 1e-5."""
    rng = np.random.default_rng(31)
    n = 3   # odd batch: the 40^2 / 20^2 maps end in a partial 128-pixel tile
    x = ((rng.integers(0, 256, (n, 3, 640, 640)).astype(np.float32)) - 127.5) / 128
    got = ctx.scrfd_forward(x)
    ref, taps = nets.scrfd_forward(det_wdict, torch.from_numpy(x), return_taps=True)
    worst = []
    for ti, t in enumerate(taps):
        t = t.numpy()
        g = ctx.scrfd_tap(ti, n, t.shape[1:])
        err = float(np.abs(g - t).max() / (np.abs(t).max() + 1e-12))
        worst.append((ti, err))
    bad = [w for w in worst if not w[1] < 2e-5]
    assert not bad, (bad[:6], worst)
    errs = []
    for i, (g, r) in enumerate(zip(got, ref)):
        r = r.numpy()
        assert g.shape == r.shape
        errs.append(float(np.abs(g - r).max()))
    print("scrfd head max abs err:", errs, "tap rel err:", worst)
    assert max(errs[:3]) < 1e-5, errs          # scores (sigmoid outputs)
    px = [e * s for e, s in zip(errs[3:], (8, 16, 32, 8, 16, 32))]
    assert max(px) < 1e-3, px                  # bbox / kps error in pixels after decode


@pytest.mark.parametrize("n", [2, 9])
def test_k2_scrfd_fused_front_matches_separate_kernels(ctx, n, monkeypatch):
    """front_fused_kernel (stem -> b0 -> s0.0 through shared-memory rings) against the three separate
    kernels: the stem and b0 use the same fragments and accumulation order, so their activations must be
    bit-identical; s0.0 moves from the tcgen05 pipeline to warp-level MMA (same three-term split), so it
    and the heads agree to fp32-grade error.  n = 2: short units (2 groups); n = 9: long units (10 groups),
    i.e. both unit lengths, every strip, frame borders on all four sides."""
    rng = np.random.default_rng(33 + n)
    x = ((rng.integers(0, 256, (n, 3, 640, 640)).astype(np.float32)) - 127.5) / 128
    x[0, :, :5] = 1.0      # something non-random on the top / left borders
    x[0, :, :, -3:] = -1.0
    monkeypatch.setenv("FR_SCRFD_FRONT_FUSED", "0")
    ref_heads = ctx.scrfd_forward(x)
    ref_taps = [ctx.scrfd_tap(i, n, s) for i, s in ((0, (16, 320, 320)), (1, (16, 320, 320)), (2, (40, 160, 160)))]
    monkeypatch.setenv("FR_SCRFD_FRONT_FUSED", "1")
    got_heads = ctx.scrfd_forward(x)
    got_taps = [ctx.scrfd_tap(i, n, s) for i, s in ((0, (16, 320, 320)), (1, (16, 320, 320)), (2, (40, 160, 160)))]
    assert np.array_equal(got_taps[0], ref_taps[0])
    assert np.array_equal(got_taps[1], ref_taps[1])
    err = float(np.abs(got_taps[2] - ref_taps[2]).max() / np.abs(ref_taps[2]).max())
    assert err < 4e-6, err
    for k, (g, r) in enumerate(zip(got_heads, ref_heads)):
        e = float(np.abs(g - r).max())
        assert e < (1e-5 if k < 3 else 2e-5), (k, e)


def test_k2_scrfd_batch_invariance(ctx):
    rng = np.random.default_rng(32)
    x = ((rng.integers(0, 256, (5, 3, 640, 640)).astype(np.float32)) - 127.5) / 128
    all5 = ctx.scrfd_forward(x)
    one = ctx.scrfd_forward(x[3:4])
    for a, b in zip(all5, one):
        assert np.array_equal(a[3], b[0])   # per-frame result independent of the batch around it
