"""1:N cosine search oracle (north_star extension; absent from the reference).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Defined as the batched
generalisation of FaceRecognizer::compareFaces (src/face_recognizer.cpp:320-334):
S = Q . G^T on L2-normalised rows, top-k per query by S (ties -> lower gallery
index), reported raw and mapped (S+1)/2; a hit is a match iff mapped > 0.6
(src/main.cpp:118-119).  PARITY UNPINNED by the reference (no such feature).
"""
from __future__ import annotations

import numpy as np


def to_bf16_f32(x: np.ndarray) -> np.ndarray:
    """Round fp32 -> bf16 (round-to-nearest-even) and widen back to fp32."""
    x = np.ascontiguousarray(x, np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(x.shape)


def topk(q: np.ndarray, g: np.ndarray, k: int, index_base: int = 0, bf16_inputs: bool = True):
    """Returns (scores [nq,k] fp32 desc, idx [nq,k] int64).  Rows padded with
    (-inf, -1) when the gallery has fewer than k rows."""
    q = np.asarray(q, np.float32)
    g = np.asarray(g, np.float32)
    if bf16_inputs:
        q, g = to_bf16_f32(q), to_bf16_f32(g)
    nq = q.shape[0]
    scores = np.full((nq, k), -np.inf, np.float32)
    idx = np.full((nq, k), -1, np.int64)
    if g.shape[0] == 0:
        return scores, idx
    s = (q.astype(np.float64) @ g.astype(np.float64).T).astype(np.float32)
    kk = min(k, g.shape[0])
    # stable sort on (-score, index): ties -> lower index first
    order = np.argsort(-s, axis=1, kind="stable")[:, :kk]
    scores[:, :kk] = np.take_along_axis(s, order, 1)
    idx[:, :kk] = order + index_base
    return scores, idx


def merge_topk(scores_parts, idx_parts, k: int):
    """Merge per-shard top-k lists (score desc, global index asc)."""
    s = np.concatenate(scores_parts, axis=1)
    i = np.concatenate(idx_parts, axis=1)
    big = np.where(i < 0, np.iinfo(np.int64).max, i)
    order = np.lexsort((big, -s.astype(np.float64)), axis=1)[:, :k]
    return np.take_along_axis(s, order, 1), np.take_along_axis(i, order, 1)


def mapped(scores: np.ndarray) -> np.ndarray:
    return ((scores.astype(np.float32) + np.float32(1.0)) / np.float32(2.0)).astype(np.float32)
