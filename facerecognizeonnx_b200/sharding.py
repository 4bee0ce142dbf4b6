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


# ------------------------------------------------------------------ enrolment / rebalancing --
# SURVEY 8(f)-4.  fr_gallery_remove fills a hole with the shard's last row, so after many removals the
# shards of a row-sharded gallery drift apart in size and the search time (max over ranks) follows the
# fullest one.  Rebalancing moves whole rows from the fullest shards to the emptiest until sizes differ
# by at most one row.  Row *identity* lives in a host-side id table per shard (ids[local_row] = person /
# record id): search returns (shard, local row) through the global index = index_base + local row, and
# the table maps that back to the id, so moving a row only has to move its table entry with it.

def rebalance_plan(sizes: List[int]) -> List[Tuple[int, int, int]]:
    """sizes[r] = rows on rank r.  Returns moves (src_rank, dst_rank, n_rows), each taking the LAST
    n_rows of src and appending them to dst, after which max(sizes) - min(sizes) <= 1.  Greedy
    largest-surplus -> largest-deficit; at most world - 1 moves."""
    world = len(sizes)
    total = sum(sizes)
    base, rem = divmod(total, world)
    # the `rem` currently-fullest ranks keep one extra row (fewest rows moved)
    order = sorted(range(world), key=lambda r: (-sizes[r], r))
    target = [base] * world
    for r in order[:rem]:
        target[r] += 1
    surplus = [[r, sizes[r] - target[r]] for r in range(world) if sizes[r] > target[r]]
    deficit = [[r, target[r] - sizes[r]] for r in range(world) if sizes[r] < target[r]]
    surplus.sort(key=lambda x: -x[1])
    deficit.sort(key=lambda x: -x[1])
    moves: List[Tuple[int, int, int]] = []
    i = j = 0
    while i < len(surplus) and j < len(deficit):
        n = min(surplus[i][1], deficit[j][1])
        moves.append((surplus[i][0], deficit[j][0], n))
        surplus[i][1] -= n
        deficit[j][1] -= n
        if surplus[i][1] == 0:
            i += 1
        if deficit[j][1] == 0:
            j += 1
    return moves


class ShardedGalleryIndex:
    """Host-side id tables of a row-sharded gallery + the add / remove / rebalance bookkeeping.
    `shards[r]` is any object with __len__, add(rows fp32 [n,512]), get_rows(first, n) and
    remove(local_row) -- capi.Gallery on a GPU, or a numpy stand-in in the CPU tests.  In a multi-process
    job every rank runs the same plan; a rank only touches its own shard and ships the moved rows
    through `transfer(src, dst, rows)` (NCCL send/recv, or a plain function call in one process)."""

    def __init__(self, shards, index_bases: List[int]):
        self.shards = shards
        self.bases = list(index_bases)
        self.ids: List[List[int]] = [[] for _ in shards]

    def sizes(self) -> List[int]:
        return [len(t) for t in self.ids]

    def add(self, rank: int, rows, ids: List[int]):
        assert len(ids) == rows.shape[0]
        self.shards[rank].add(rows)
        self.ids[rank].extend(int(i) for i in ids)

    def remove_id(self, record_id: int) -> bool:
        """Removes one record; the shard's last row moves into its place (fr_gallery_remove)."""
        for r, table in enumerate(self.ids):
            if record_id in table:
                row = table.index(record_id)
                self.shards[r].remove(row)
                table[row] = table[-1]
                table.pop()
                return True
        return False

    def resolve(self, global_index: int) -> int:
        """global row index (as returned by search) -> record id; -1 for an empty slot."""
        if global_index < 0:
            return -1
        for r in range(len(self.shards) - 1, -1, -1):
            if global_index >= self.bases[r]:
                local = global_index - self.bases[r]
                return self.ids[r][local] if local < len(self.ids[r]) else -1
        return -1

    def rebalance(self, transfer=None) -> List[Tuple[int, int, int]]:
        """Executes rebalance_plan.  Rows travel as the fp32 image of the stored bf16 values, so the
        copy is bit-exact (bf16 -> fp32 -> bf16 is the identity)."""
        moves = rebalance_plan(self.sizes())
        for src, dst, n in moves:
            first = len(self.ids[src]) - n
            rows = self.shards[src].get_rows(first, n)
            if transfer is not None:
                rows = transfer(src, dst, rows)
            moved_ids = self.ids[src][first:]
            for _ in range(n):                       # drop the tail of src (last row first: no data movement)
                self.shards[src].remove(len(self.ids[src]) - 1)
                self.ids[src].pop()
            self.shards[dst].add(rows)
            self.ids[dst].extend(moved_ids)
        return moves
