"""GPU parity for K9 (1:N cosine search with fused top-k) against the numpy oracle.

Scores come from bf16 x bf16 products accumulated in fp32 on the tensor core; the oracle uses
the same bf16-rounded inputs with a float64 accumulate, so scores agree to ~1e-5 and indices
agree wherever neighbouring scores are further apart than that."""
import numpy as np
import pytest

from oracle import gallery as ogal

pytestmark = pytest.mark.gpu


def _unit(rng, n):
    x = rng.normal(size=(n, 512)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def _check(s, i, g, q, k, base=0):
    ref_s, ref_i = ogal.topk(q, g, k, index_base=base)
    assert np.allclose(s, ref_s, atol=2e-5), float(np.abs(s - ref_s).max())
    gap_ok = np.ones_like(ref_i, bool)
    d = np.abs(np.diff(ref_s, axis=1)) > 1e-4
    gap_ok[:, 1:] &= d
    gap_ok[:, :-1] &= d
    assert np.array_equal(i[gap_ok], ref_i[gap_ok])
    # every returned index must carry its own score
    gq = ogal.to_bf16_f32(q).astype(np.float64)
    gg = ogal.to_bf16_f32(g).astype(np.float64)
    for r in range(q.shape[0]):
        ok = i[r] >= 0
        assert np.allclose((gq[r] @ gg[i[r][ok] - base].T), s[r][ok], atol=2e-5)


@pytest.mark.parametrize("n_rows,nq,k", [(5000, 200, 10), (256, 128, 10), (1000, 1, 5), (70000, 300, 16), (7, 3, 10)])
def test_k9_search_vs_oracle(ctx, capi, n_rows, nq, k):
    rng = np.random.default_rng(n_rows + nq)
    g = _unit(rng, n_rows)
    q = _unit(rng, nq)
    q[: min(nq, 50)] = g[rng.integers(0, n_rows, min(nq, 50))]   # planted exact matches
    gal = capi.Gallery(ctx, n_rows + 100, index_base=1000)
    gal.add(g[: n_rows // 2])
    gal.add(g[n_rows // 2:])
    assert len(gal) == n_rows
    s, i = gal.search(q, k)
    if n_rows < k:
        assert np.all(i[:, n_rows:] == -1) and np.all(np.isinf(s[:, n_rows:]))
        s, i, k = s[:, :n_rows], i[:, :n_rows], n_rows
    _check(s, i, g, q, k, base=1000)
    gal.close()


def test_k9_ties_prefer_lower_index(ctx, capi):
    rng = np.random.default_rng(3)
    g = _unit(rng, 3000)
    g[2500] = g[10]
    g[700] = g[10]
    gal = capi.Gallery(ctx, 3000)
    gal.add(g)
    s, i = gal.search(g[10:11], 5)
    assert list(i[0][:3]) == [10, 700, 2500]
    gal.close()


def test_k9_synthetic_gallery_and_planted_queries(ctx, capi):
    gal = capi.Gallery(ctx, 200000, index_base=0)
    gal.fill_synthetic(200000, seed=1000)
    rows = gal.get_rows(0, 200000)
    assert np.allclose(np.linalg.norm(rows, axis=1), 1.0, atol=1e-2)
    rng = np.random.default_rng(4)
    planted = rng.integers(0, 200000, 256)
    q = rows[planted] + 0.02 * rng.normal(size=(256, 512)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    s, i = gal.search(q, 10)
    assert np.array_equal(i[:, 0], planted)
    _check(s, i, rows, q, 10)
    gal.close()


def test_k9_sharded_merge_equals_single(ctx, capi):
    """Two shards on one GPU + fr_topk_merge == one gallery (the N-rank path minus NCCL)."""
    rng = np.random.default_rng(5)
    g = _unit(rng, 4001)
    q = _unit(rng, 130)
    whole = capi.Gallery(ctx, 4001)
    whole.add(g)
    ws, wi = whole.search(q, 10)
    parts_s, parts_i = [], []
    from facerecognizeonnx_b200 import sharding
    for r in range(3):
        lo, hi = sharding.shard_range(4001, r, 3)
        sh = capi.Gallery(ctx, hi - lo, index_base=lo)
        sh.add(g[lo:hi])
        s, i = sh.search(q, 10)
        parts_s.append(s)
        parts_i.append(i)
        sh.close()
    ms, mi = capi.topk_merge(ctx, np.stack(parts_s), np.stack(parts_i), 10)
    assert np.array_equal(mi, wi) and np.array_equal(ms, ws)
    os_, oi = ogal.merge_topk(parts_s, parts_i, 10)
    assert np.array_equal(mi, oi)
    whole.close()


def test_gallery_save_load_remove_roundtrip(ctx, capi, tmp_path):
    """a saved shard reloads bit for bit (same search results), rows can be
    appended from several files, and remove() keeps the result set consistent."""
    rng = np.random.default_rng(5)
    g = rng.normal(size=(3000, 512)).astype(np.float32)
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    q = g[rng.integers(0, 3000, 40)] + 0.02 * rng.normal(size=(40, 512)).astype(np.float32)
    a = capi.Gallery(ctx, 4096, index_base=7000)
    a.add(g[:2000])
    p1, p2 = str(tmp_path / "a.frg"), str(tmp_path / "b.frg")
    a.save(p1)
    b = capi.Gallery(ctx, 1000, index_base=0)
    b.add(g[2000:])
    b.save(p2)
    c = capi.Gallery(ctx, 4096, index_base=7000)
    assert c.load(p1) == 7000 and c.load(p2) == 0 and len(c) == 3000
    assert np.array_equal(c.get_rows(0, 3000), np.concatenate([a.get_rows(0, 2000), b.get_rows(0, 1000)]))
    a.add(g[2000:])
    s_a, i_a = a.search(q, 10)
    s_c, i_c = c.search(q, 10)
    assert np.array_equal(s_a, s_c) and np.array_equal(i_a, i_c)
    # remove the best match of query 0: it must disappear, the former last row takes its index
    top = int(i_c[0, 0]) - 7000
    c.remove(top)
    assert len(c) == 2999
    s2, i2 = c.search(q[:1], 10)
    ref = np.delete(np.arange(3000), top)
    g_bf = c.get_rows(0, 2999)
    assert np.array_equal(g_bf[top], a.get_rows(2999, 1)[0])
    exp = np.sort((g_bf @ q[0]).astype(np.float32))[::-1][:10]
    assert np.allclose(s2[0], exp, atol=3e-3)
    with pytest.raises(capi.FrError):
        capi.Gallery(ctx, 10).load(p1)            # does not fit
    bad = str(tmp_path / "bad.frg")
    open(bad, "wb").write(b"not a gallery")
    with pytest.raises(capi.FrError):
        c.load(bad)


def test_remove_then_rebalance_across_shards_keeps_results(ctx, capi):
    """enrolment with removal + shard rebalancing on real galleries (four shards on one
    GPU standing in for four ranks): after the plan runs, sizes differ by <= 1 row, every surviving
    record is its own top-1 under its id, moved rows are bit-identical, removed ids are gone."""
    from facerecognizeonnx_b200 import sharding
    rng = np.random.default_rng(11)
    world, per = 4, 300
    rows = rng.normal(size=(world * per, 512)).astype(np.float32)
    rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    bases = [r * 100000 for r in range(world)]
    shards = [capi.Gallery(ctx, 2 * per, index_base=b) for b in bases]
    idx = sharding.ShardedGalleryIndex(shards, bases)
    for r in range(world):
        idx.add(r, rows[r * per:(r + 1) * per], list(range(r * per, (r + 1) * per)))
    removed = set(int(i) for i in rng.choice(np.arange(0, per), 250, replace=False)) | {per + 7, 3 * per + 1}
    for rid in removed:
        assert idx.remove_id(rid)
    before = {rid: shards[r].get_rows(j, 1)[0] for r, t in enumerate(idx.ids) for j, rid in enumerate(t)}
    assert [len(s) for s in shards] == idx.sizes() == [50, 299, 300, 299]
    moves = idx.rebalance()
    assert moves and max(idx.sizes()) - min(idx.sizes()) <= 1 and [len(s) for s in shards] == idx.sizes()
    after = {rid: shards[r].get_rows(j, 1)[0] for r, t in enumerate(idx.ids) for j, rid in enumerate(t)}
    assert set(after) == set(before) and all(np.array_equal(after[k], before[k]) for k in after)   # bit-exact moves
    alive = sorted(after)
    k = 3
    parts = [s.search(rows[alive], k) for s in shards]
    ms, mi = capi.topk_merge(ctx, np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]), k)
    assert [idx.resolve(int(g)) for g in mi[:, 0]] == alive and np.all(ms[:, 0] > 0.99)
    gone = sorted(removed)
    parts = [s.search(rows[gone], 1) for s in shards]
    ms, mi = capi.topk_merge(ctx, np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]), 1)
    assert all(idx.resolve(int(g)) not in removed for g in mi[:, 0]) and np.all(ms[:, 0] < 0.5)
    for s in shards:
        s.close()


def _planted(rng, n_rows, n_planted_q, k):
    """Random unit gallery; for the first n_planted_q queries, k rows planted at cosine 0.9, 0.85, ..."""
    rows = rng.normal(size=(n_rows, 512)).astype(np.float32)
    rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    q = rng.normal(size=(n_planted_q + 96, 512)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    slots = rng.choice(n_rows, n_planted_q * k, replace=False).reshape(n_planted_q, k)
    for qi in range(n_planted_q):
        for j in range(k):
            a = 0.9 - 0.05 * j
            noise = rng.normal(size=512).astype(np.float32)
            noise -= noise.dot(q[qi]) * q[qi]
            noise /= np.linalg.norm(noise)
            rows[slots[qi, j]] = a * q[qi] + np.sqrt(1 - a * a) * noise
    return rows, q, slots


def test_fp8_gallery_coarse_pass_plus_bf16_rerank(ctx, capi):
    """e4m3 gallery (kind::f8f6f4 coarse pass) with exact bf16 re-rank.
    * every returned score is the bf16 search's score of that row (|d| <= 2e-6: summation order only);
    * the re-ranked top-10 equals the bf16 search wherever the oracle's score gap between rank k and the
      first candidate that could be lost (rank 17) exceeds twice the fp8 error bound;
    * on random queries (gaps inside the fp8 noise) recall stays >= 0.9."""
    from oracle import gallery as ogal
    rng = np.random.default_rng(21)
    n_rows, nq_p, k = 70_001, 64, 10
    rows, q, slots = _planted(rng, n_rows, nq_p, k)
    g = capi.Gallery(ctx, n_rows, index_base=500, flags=capi.Gallery.FP8)
    for first in range(0, n_rows, 20_000):
        g.add(rows[first:first + 20_000])
    s_b, i_b = g.search(q, k)
    s_f, i_f = g.search_fp8(q, k)
    assert np.array_equal(i_f[:nq_p], i_b[:nq_p]) and np.array_equal(i_b[:nq_p] - 500, slots)   # planted: exact
    # re-rank exactness: the score of every returned row, recomputed by the numpy oracle from bf16 values
    qb, rb = ogal.to_bf16_f32(q), ogal.to_bf16_f32(rows)
    exact = np.einsum("qd,qkd->qk", qb.astype(np.float64), rb[i_f - 500].astype(np.float64))
    assert np.abs(s_f - exact).max() < 2e-6, float(np.abs(s_f - exact).max())
    assert np.all(np.diff(s_f, axis=1) <= 0)
    # gap rule on all queries, with the bf16 search's own top-16 as the oracle list
    s16, _ = g.search(q, 16)
    eps = 0.02                                   # 5 sigma of the e4m3 score error (sigma ~ 4e-3)
    safe = (s16[:, k - 1] - s16[:, 15]) > 2 * eps
    assert safe[:nq_p].all()
    assert np.array_equal(i_f[safe], i_b[safe])
    recall = np.mean([len(set(a) & set(b)) / k for a, b in zip(i_f[nq_p:], i_b[nq_p:])])
    assert recall >= 0.9, recall
    # small batches are spread over > 8 splits: the group-merge path must give the same answer
    s1, i1 = g.search_fp8(q[:3], k)
    assert np.array_equal(i1, i_f[:3]) and np.abs(s1 - s_f[:3]).max() < 2e-6
    g.close()


def test_fp8_gallery_host_resident_bf16_rows_and_persistence(ctx, capi, tmp_path):
    """Capacity mode: only the e4m3 rows in HBM, bf16 rows in mapped pinned host memory.  Same results as
    the all-HBM fp8 gallery; the plain bf16 search refuses; save / load / remove keep both copies in step."""
    rng = np.random.default_rng(22)
    n_rows, k = 20_000, 10
    rows, q, _ = _planted(rng, n_rows, 16, k)
    dev = capi.Gallery(ctx, n_rows, flags=capi.Gallery.FP8)
    host = capi.Gallery(ctx, n_rows, flags=capi.Gallery.FP8 | capi.Gallery.BF16_ON_HOST)
    dev.add(rows)
    host.add(rows)
    sd, idd = dev.search_fp8(q, k)
    sh, ih = host.search_fp8(q, k)
    assert np.array_equal(idd, ih) and np.array_equal(sd, sh)
    with pytest.raises(capi.FrError) as e:
        host.search(q, k)
    assert e.value.code == capi.FR_ERR_UNSUPPORTED
    with pytest.raises(capi.FrError):
        capi.Gallery(ctx, 10, flags=capi.Gallery.BF16_ON_HOST)          # needs FP8
    plain = capi.Gallery(ctx, 10)
    with pytest.raises(capi.FrError):
        plain.search_fp8(q[:1], 1)                                      # no e4m3 mirror
    assert np.array_equal(host.get_rows(0, 50), dev.get_rows(0, 50))
    # persistence: the file holds the bf16 rows; the e4m3 mirror is rebuilt on load
    p = str(tmp_path / "shard.frg")
    host.save(p)
    again = capi.Gallery(ctx, n_rows, flags=capi.Gallery.FP8)
    again.load(p)
    s2, i2 = again.search_fp8(q, k)
    assert np.array_equal(i2, idd) and np.array_equal(s2, sd)
    # remove: the last row moves into the hole in both copies
    victim = int(idd[0, 0])
    for gal in (dev, host):
        gal.remove(victim)
    s3, i3 = dev.search_fp8(q[:1], k)
    s4, i4 = host.search_fp8(q[:1], k)
    assert np.array_equal(i3, i4) and np.array_equal(s3, s4) and victim not in i3[0][s3[0] > 0.8]
    moved = dev.get_rows(victim, 1)[0]
    assert np.array_equal(moved, __import__("oracle.gallery", fromlist=["x"]).to_bf16_f32(rows[n_rows - 1]))
    for gal in (dev, host, again, plain):
        gal.close()
