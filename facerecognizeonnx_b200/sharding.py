"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed plumbing).

* Frame batches (det + align + embed) are independent units (the reference's detect /
  extractFeature are pure functions of their arguments, src/face_detector.cpp:139-222,
  src/face_recognizer.cpp:236-304): data-parallel, NO data-path collective.
* The 1:N gallery is row-sharded: rank r owns global rows [lo_r, hi_r); every rank searches
  its shard with global indices, the per-rank top-k lists are all-gathered (NCCL over NVLink on
  GPUs, gloo in the CPU tests) and merged with the (score desc, global index asc) order, so
  the result is independent of the number of ranks.

The functions take the local search / merge callables as arguments so that the same plumbing
runs with the CUDA library on GPUs and with the numpy oracle in the world_size-2 gloo tests.
"""
from __future__ import annotations

from typing import Callable, List, Tuple


def shard_range(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row range [lo, hi) of `rank`; sizes differ by at most one row."""
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def frames_for_rank(n_frames: int, rank: int, world: int) -> range:
    """Contiguous chunk of frame indices processed by `rank` (data parallel)."""
    lo, hi = shard_range(n_frames, rank, world)
    return range(lo, hi)


def sharded_search(local_search: Callable, merge: Callable, queries, k: int, group=None):
    """local_search(queries, k) -> (scores [nq,k] tensor, idx [nq,k] int64 tensor, global
    indices).  merge(scores [W,nq,k], idx [W,nq,k], k) -> merged pair.  Returns the merged
    top-k, identical on every rank."""
    import torch
    import torch.distributed as dist

    s, i = local_search(queries, k)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return merge(s.unsqueeze(0), i.unsqueeze(0), k)
    nq = s.shape[0]
    gs = torch.empty((world * nq,) + tuple(s.shape[1:]), dtype=s.dtype, device=s.device)
    gi = torch.empty((world * nq,) + tuple(i.shape[1:]), dtype=i.dtype, device=i.device)
    dist.all_gather_into_tensor(gs, s.contiguous(), group=group)   # rank-major concatenation
    dist.all_gather_into_tensor(gi, i.contiguous(), group=group)
    return merge(gs.view(world, nq, -1), gi.view(world, nq, -1), k)


def sharded_search_packed(local_search_packed: Callable, merge_packed: Callable, queries, k: int, group=None):
    """The same exchange with ONE collective: local_search_packed(queries, k) -> int64 tensor [nq, k]
    of packed records (low word fp32 score bits, high word global row index; include/fr_capi.h
    fr_gallery_search_packed), all-gathered rank-major, then merge_packed(records [W, nq, k], k)."""
    import torch
    import torch.distributed as dist

    rec = local_search_packed(queries, k)
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return merge_packed(rec.unsqueeze(0), k)
    nq = rec.shape[0]
    gathered = torch.empty((world * nq, rec.shape[1]), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(gathered, rec.contiguous(), group=group)
    return merge_packed(gathered.view(world, nq, -1), k)
