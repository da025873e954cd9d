"""Multi-GPU plumbing for the tile-row partitioned SpGEMM (one process per GPU, torch.distributed).

The path shards by C tile-rows: C tile-row I depends only on A tile-row I and on B.
  * rank 0 tiles B once on its GPU and broadcasts it as ONE buffer over NCCL/NVLink (the tiled
    layout is a pure function of the sizes, tsg_tile_alloc);
  * A's tile-rows are cut into `world` contiguous ranges balanced by the step-1 weight
    w[I] = number of matched tile pairs of tile-row I (tsg_tilerow_weights); each rank receives the
    CSR rows of its range and tiles them locally;
  * every rank runs steps 1-3 on its range with no further communication;
  * C stays distributed; "concatenation" is an all-gather of the per-rank (tiles, nnz) counts that
    rebases tile_ptr / tile_nnz offsets in 64 bits.
The host-side logic here (partitioning, offset rebasing, the size exchange) is backend-agnostic and is
covered by world_size-2 gloo tests on CPU; only the payload broadcasts need NCCL.
"""
from __future__ import annotations

import numpy as np


def partition_tilerows(weights: np.ndarray, parts: int, min_weight: float = 1.0) -> np.ndarray:
    """Cut tile-rows [0, len(weights)) into `parts` contiguous ranges of near-equal total weight.

    Returns `parts + 1` ascending cut points (first 0, last len(weights)). Every tile-row costs at
    least `min_weight` so that empty tile-rows are still spread. Cuts are placed where the prefix sum
    crosses k * total / parts."""
    w = np.maximum(np.asarray(weights, dtype=np.float64), min_weight)
    n = w.size
    prefix = np.concatenate([[0.0], np.cumsum(w)])
    total = prefix[-1]
    cuts = [0]
    for k in range(1, parts):
        target = total * k / parts
        c = int(np.searchsorted(prefix, target, side="left"))
        # choose the nearer of the two neighbouring cut positions
        if c > 0 and abs(prefix[c - 1] - target) <= abs(prefix[min(c, n)] - target):
            c -= 1
        c = max(c, cuts[-1])
        cuts.append(min(c, n))
    cuts.append(n)
    return np.asarray(cuts, dtype=np.int64)


def imbalance(weights: np.ndarray, cuts: np.ndarray) -> float:
    """max part weight / mean part weight (1.0 = perfectly balanced)."""
    w = np.asarray(weights, dtype=np.float64)
    sums = np.array([w[a:b].sum() for a, b in zip(cuts[:-1], cuts[1:])])
    return float(sums.max() / max(sums.mean(), 1e-300))


def concat_offsets(counts: np.ndarray) -> np.ndarray:
    """Exclusive 64-bit offsets of per-rank counts (rows: ranks; columns: quantities)."""
    c = np.asarray(counts, dtype=np.int64)
    out = np.zeros_like(c)
    out[1:] = np.cumsum(c[:-1], axis=0)
    return out


def gather_counts(local_counts, dist=None) -> np.ndarray:
    """All-gather a small vector of int64 counts; returns array [world, len]. Works on any backend
    (tensors are moved to the GPU for NCCL)."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray([local_counts], dtype=np.int64)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(list(local_counts), dtype=torch.int64, device=dev)
    outs = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return np.stack([o.cpu().numpy() for o in outs])


def csr_row_slice(rowptr: np.ndarray, colidx: np.ndarray, val: np.ndarray, r0: int, r1: int):
    """Rows [r0, r1) of a CSR matrix as a CSR of its own (row pointer rebased)."""
    lo, hi = int(rowptr[r0]), int(rowptr[r1])
    return (rowptr[r0:r1 + 1] - rowptr[r0]).astype(np.int32), colidx[lo:hi], val[lo:hi]


class DeviceBuffer:
    """Zero-copy view of library-owned device memory for torch (``torch.as_tensor(buf, device=...)``)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def tile_slab_tensor(tile, device):
    """The single device slab holding a tiled matrix (slab[0]) as a uint8 torch tensor."""
    import torch
    return torch.as_tensor(DeviceBuffer(tile.d.slab[0], tile.d.slab_bytes[0]), device=device)
