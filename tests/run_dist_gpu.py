"""Multi-GPU check, launched with torchrun (one process per GPU), not collected by pytest:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/run_dist_gpu.py

* row-sharded 1:N search + NCCL all-gather + fr_topk_merge == numpy oracle on the whole gallery
  == the same search on one GPU (rank-count invariance);
* data-parallel det+align+embed: every rank's per-frame results equal a single-GPU run of the
  same frames (no cross-frame state, no collective).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from facerecognizeonnx_b200 import capi, sharding  # noqa: E402
from oracle import gallery as ogal  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = capi.Context(local, capi.Weights(capi.FR_MODEL_DET, None, 1), capi.Weights(capi.FR_MODEL_REC, None, 1))
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    # ---- sharded gallery
    rng = np.random.default_rng(0)
    n_rows, nq, k = 30011, 257, 10
    g = rng.normal(size=(n_rows, 512)).astype(np.float32)
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    g[n_rows - 5] = g[3]                       # duplicate rows in different shards
    q = g[rng.integers(0, n_rows, nq)] + 0.05 * rng.normal(size=(nq, 512)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[0] = g[3]
    lo, hi = sharding.shard_range(n_rows, rank, world)
    gal = capi.Gallery(ctx, hi - lo, index_base=lo)
    gal.add(g[lo:hi])
    qd = torch.from_numpy(q).to(dev)

    def local_search(queries, kk):
        s = torch.empty(nq, kk, dtype=torch.float32, device=dev)
        i = torch.empty(nq, kk, dtype=torch.int64, device=dev)
        gal.search_dev(queries.data_ptr(), nq, kk, s.data_ptr(), i.data_ptr())
        return s, i

    def merge(gs, gi, kk):
        s = torch.empty(nq, kk, dtype=torch.float32, device=dev)
        i = torch.empty(nq, kk, dtype=torch.int64, device=dev)
        capi.topk_merge_dev(ctx, gs.contiguous().data_ptr(), gi.contiguous().data_ptr(), gs.shape[0], nq, kk,
                            s.data_ptr(), i.data_ptr())
        return s, i

    s, i = sharding.sharded_search(local_search, merge, qd, k)
    torch.cuda.synchronize()

    # the single-all-gather exchange (packed 8-byte records) must give the same answer
    def local_packed(queries, kk):
        rec = torch.empty(nq, kk, dtype=torch.int64, device=dev)
        gal.search_packed_dev(queries.data_ptr(), nq, kk, rec.data_ptr())
        return rec

    def merge_packed(rec, kk):
        ss = torch.empty(nq, kk, dtype=torch.float32, device=dev)
        ii = torch.empty(nq, kk, dtype=torch.int64, device=dev)
        capi.topk_merge_packed_dev(ctx, rec.contiguous().data_ptr(), rec.shape[0], nq, kk, ss.data_ptr(), ii.data_ptr())
        return ss, ii

    s2, i2 = sharding.sharded_search_packed(local_packed, merge_packed, qd, k)
    torch.cuda.synchronize()
    assert torch.equal(s, s2) and torch.equal(i, i2)
    s, i = s.cpu().numpy(), i.cpu().numpy()
    ref_s, ref_i = ogal.topk(q, g, k)
    assert np.allclose(s, ref_s, atol=2e-5), float(np.abs(s - ref_s).max())
    gap = np.ones_like(ref_i, bool)
    d = np.abs(np.diff(ref_s, axis=1)) > 1e-4
    gap[:, 1:] &= d
    gap[:, :-1] &= d
    assert np.array_equal(i[gap], ref_i[gap])
    assert list(i[0][:2]) == [3, n_rows - 5], i[0][:3]     # tie -> lower global index, across shards
    # identical on every rank
    ti = torch.from_numpy(i).to(dev)
    t0 = ti.clone()
    dist.broadcast(t0, 0)
    assert torch.equal(ti, t0)

    # ---- data-parallel frames: rank r processes its chunk; compare with rank 0 doing all frames
    frames = [np.random.default_rng(100 + f).integers(0, 256, (640, 640, 3), dtype=np.uint8) for f in range(4 * world)]
    mine = sharding.frames_for_rank(len(frames), rank, world)
    dets = ctx.detect_batch([frames[f] for f in mine], 0.5, 0.4, cap=64)
    counts = torch.tensor([len(d) for d in dets], device=dev)
    allc = torch.empty(world * len(mine), dtype=counts.dtype, device=dev)
    dist.all_gather_into_tensor(allc, counts)
    if rank == 0:
        full = ctx.detect_batch(frames, 0.5, 0.4, cap=64)
        assert [len(d) for d in full] == allc.cpu().tolist()
        for f in mine:
            assert np.array_equal(full[f], dets[f - mine[0]])
    dist.barrier()
    if rank == 0:
        print(f"run_dist_gpu ok: world={world} sharded search + DP detect verified", flush=True)
    gal.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
