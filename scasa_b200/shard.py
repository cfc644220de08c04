"""Multi-GPU sharding of the quant path: by CHUNK, never by cell.

(chunk, block) tasks are independent, but the cells of one chunk are coupled through estim.X.em
and the chunk-wide convergence tests (Scasa_Rsource.R:755-756, :1019-1020), so a chunk is the
unit that may move between GPUs.  Chunk boundaries are those of the whole sample
(scasa_quant_step2_v1.0.0.R:64-68); every rank takes a contiguous range of chunks, balanced by
work.  There is no collective on the data path: ranks write disjoint row ranges of the result.
Only the row filter of step2:167 needs a cross-rank value — the per-column sums — and those are
added on the host in rank order (a fixed order, so the result does not depend on timing).
"""
from __future__ import annotations

import numpy as np


def shard_chunks(chunk_off, world: int, weights=None):
    """Contiguous chunk ranges per rank: list of (first_chunk, last_chunk_exclusive).

    weights: per-chunk work estimate (default: cells per chunk).  Greedy prefix split at the
    points closest to k/world of the total weight; every rank gets at least one chunk while
    chunks last."""
    chunk_off = np.asarray(chunk_off, np.int64)
    n_chunks = len(chunk_off) - 1
    w = np.diff(chunk_off).astype(np.float64) if weights is None else np.asarray(weights, np.float64)
    assert len(w) == n_chunks
    cum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for k in range(1, world):
        target = cum[-1] * k / world
        c = int(np.searchsorted(cum, target, side="left"))
        if c > 0 and abs(cum[c - 1] - target) <= abs(cum[min(c, n_chunks)] - target):
            c -= 1
        c = max(c, cuts[-1] + 1) if cuts[-1] + 1 <= n_chunks else n_chunks
        cuts.append(min(c, n_chunks))
    cuts.append(n_chunks)
    return [(cuts[r], max(cuts[r], cuts[r + 1])) for r in range(world)]


def local_problem(chunk_off, cell_ptr, rank_range):
    """Cells and entry range of one rank: (cell0, cell1, local chunk_off, local cell_ptr slice bounds)."""
    c0, c1 = rank_range
    chunk_off = np.asarray(chunk_off, np.int64)
    cell0, cell1 = int(chunk_off[c0]), int(chunk_off[c1])
    local_chunks = chunk_off[c0:c1 + 1] - cell0
    e0, e1 = int(cell_ptr[cell0]), int(cell_ptr[cell1])
    return cell0, cell1, local_chunks, e0, e1


def combine_colsum(local_colsum: np.ndarray, group=None) -> np.ndarray:
    """Column sums over all ranks, added in rank order on the host (gloo or nccl all_gather of the
    small per-rank vectors; 33k doubles)."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return np.asarray(local_colsum, np.float64)
    world = dist.get_world_size(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.as_tensor(np.asarray(local_colsum, np.float64), device=dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    out = np.zeros(len(local_colsum), np.float64)
    for p in parts:                       # fixed order
        out += p.cpu().numpy()
    return out
