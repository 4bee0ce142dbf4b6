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
