"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard ranges, data-parallel frame
assignment and the all-gather + merge of the row-sharded 1:N search, with the numpy oracle as
the local search / merge."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from facerecognizeonnx_b200 import sharding
from oracle import gallery as ogal


def test_shard_ranges_cover_and_balance():
    for n, w in [(10_000_000, 8), (1000, 3), (7, 8), (0, 2)]:
        rs = [sharding.shard_range(n, r, w) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == n
        assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in rs]
        assert max(sizes) - min(sizes) <= 1
    assert [list(sharding.frames_for_rank(5, r, 2)) for r in range(2)] == [[0, 1, 2], [3, 4]]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, g, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.shard_range(g.shape[0], rank, world)

    def local_search(queries, kk):
        s, i = ogal.topk(queries, g[lo:hi], kk, index_base=lo)
        return torch.from_numpy(s), torch.from_numpy(i)

    def merge(gs, gi, kk):
        s, i = ogal.merge_topk(list(gs.numpy()), list(gi.numpy()), kk)
        return torch.from_numpy(s), torch.from_numpy(i)

    s, i = sharding.sharded_search(local_search, merge, q, k)
    out[rank] = (s.numpy(), i.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_search_is_rank_count_invariant():
    rng = np.random.default_rng(0)
    g = rng.normal(size=(1001, 512)).astype(np.float32)
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    g[500] = g[20]            # exact duplicate rows in different shards: tie -> lower global index
    g[900] = g[20]
    q = g[[20, 7, 999]] + 0.01 * rng.normal(size=(3, 512)).astype(np.float32)
    q[0] = g[20]
    k = 10
    ref_s, ref_i = ogal.topk(q, g, k)
    assert list(ref_i[0][:3]) == [20, 500, 900]
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, q, g, k, out), nprocs=2, join=True)
    for r in range(2):
        s, i = out[r]
        assert np.array_equal(i, ref_i)
        assert np.allclose(s, ref_s, atol=1e-6)


# ------------------------------------------------------------------ rebalancing --
class _NumpyShard:
    """CPU stand-in for capi.Gallery with the same add / get_rows / remove semantics (bf16 storage,
    remove = move the last row into the hole)."""

    def __init__(self):
        self.rows = np.zeros((0, 512), np.float32)

    def __len__(self):
        return self.rows.shape[0]

    def add(self, rows):
        self.rows = np.concatenate([self.rows, ogal.to_bf16_f32(rows)])

    def get_rows(self, first, n):
        return self.rows[first:first + n].copy()

    def remove(self, row):
        self.rows[row] = self.rows[-1]
        self.rows = self.rows[:-1]


def test_rebalance_plan_properties():
    rng = np.random.default_rng(0)
    for _ in range(200):
        w = int(rng.integers(1, 9))
        sizes = [int(x) for x in rng.integers(0, 1000, w)]
        moves = sharding.rebalance_plan(sizes)
        after = list(sizes)
        for s, d, n in moves:
            assert n > 0 and s != d and after[s] >= n
            after[s] -= n
            after[d] += n
        assert sum(after) == sum(sizes) and max(after) - min(after) <= 1
        assert len(moves) <= max(w - 1, 0)
        # minimal traffic: exactly the surplus over the balanced sizes moves
        assert sum(n for _, _, n in moves) == sum(max(0, a - b) for a, b in zip(sizes, after))
    assert sharding.rebalance_plan([5, 5, 5]) == []


def test_sharded_index_remove_and_rebalance_keep_search_results():
    """Enrol -> remove many records from one shard -> rebalance: shard sizes even out, every surviving
    record is still found as its own top-1 with the right id, removed ids never come back."""
    rng = np.random.default_rng(1)
    world, per = 4, 120
    rows = rng.normal(size=(world * per, 512)).astype(np.float32)
    rows /= np.linalg.norm(rows, axis=1, keepdims=True)
    bases = [r * 1000 for r in range(world)]
    idx = sharding.ShardedGalleryIndex([_NumpyShard() for _ in range(world)], bases)
    for r in range(world):
        idx.add(r, rows[r * per:(r + 1) * per], list(range(r * per, (r + 1) * per)))
    removed = [int(i) for i in rng.choice(np.arange(0, per), 90, replace=False)] + [130, 250]   # mostly shard 0
    for rid in removed:
        assert idx.remove_id(rid)
    assert not idx.remove_id(removed[0])
    assert idx.sizes() == [30, 119, 119, 120]
    moves = idx.rebalance()
    assert moves and max(idx.sizes()) - min(idx.sizes()) <= 1 and sum(idx.sizes()) == world * per - 92

    def search(q, k):
        parts = [ogal.topk(q, s.rows, k, index_base=b) for s, b in zip(idx.shards, bases)]
        return ogal.merge_topk([p[0] for p in parts], [p[1] for p in parts], k)

    alive = [i for i in range(world * per) if i not in set(removed)]
    s, gi = search(rows[alive], 3)
    found = [idx.resolve(int(g)) for g in gi[:, 0]]
    assert found == alive and np.all(s[:, 0] > 0.99)
    s, gi = search(rows[removed], 1)
    assert all(idx.resolve(int(g)) not in set(removed) for g in gi[:, 0]) and np.all(s[:, 0] < 0.5)
