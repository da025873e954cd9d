"""Multi-GPU tile-row partitioned SpGEMM (one process per GPU, torch.distributed over NCCL / NVLink).

The path shards by C tile-rows: C tile-row I depends only on A tile-row I and on B, so there is one
exchange step and no collective on the compute path. The reference has nothing here (src/main.cu:41-49 picks one device).

    shard = multigpu.distribute(A_host_on_rank0, aat, dist, device)   # collective: every rank calls it
    totals, per_slab = multigpu.spgemm(shard, sink=...)               # local: this rank's C tile-rows
    offsets = multigpu.concat(shard, totals)                          # 64-bit offsets of this rank's C inside the whole C
    csr = multigpu.gather_csr(shard, local_csr)                       # optional: whole CSR(C) on rank 0
or, in one call (SURVEY.md 8b's additive `tilespgemm_multi`):
    res = multigpu.run(A_host_on_rank0, aat, dist, device, gather=True)

distribute():
  * B is broadcast ONCE over NCCL -- as the CSR it is built from (mode "csr", default: for C = A^2 that CSR is A itself
    and for C = A A^T it is A, transposed on every GPU; 12 bytes per nonzero, where the tiled form of a hypersparse
    matrix is 5-7x larger) or as the tiled matrix in one buffer (mode "tiled": the tiled layout is a
    pure function of the sizes, tsg_tile_alloc). Every rank then holds the whole tiled B.
  * the step-1 weights w[I] (matched tile pairs of tile-row I, tsg_tilerow_weights) are computed in parallel -- every rank
    tiles an equal block of A's rows straight out of the broadcast CSR (tsg_csr_row_slice: no host round trip) -- and
    all-gathered; the tile-rows are cut into `world` contiguous ranges of equal weight; each rank tiles its range.
  * nothing is sliced on the host and nothing is sent point to point.
The host-side logic (partitioning, weight exchange, offset rebasing) is backend-agnostic and is covered by world_size-2
gloo tests on CPU; the payload collectives and everything that touches the library need a GPU.
"""
from __future__ import annotations

from dataclasses import dataclass, field

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


def _world(dist):
    if dist is None or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(), dist.get_world_size()


def _comm_device(dist, device):
    return device if dist is not None and dist.is_initialized() and dist.get_backend() == "nccl" else "cpu"


def gather_counts(local_counts, dist=None, device="cuda") -> np.ndarray:
    """All-gather a small vector of int64 counts; returns array [world, len]. Works on any backend
    (tensors live on the GPU for NCCL)."""
    import torch
    rank, world = _world(dist)
    if world == 1:
        return np.asarray([local_counts], dtype=np.int64)
    t = torch.tensor([int(x) for x in local_counts], dtype=torch.int64, device=_comm_device(dist, device))
    outs = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return np.stack([o.cpu().numpy() for o in outs])


def equal_row_blocks(tilem: int, world: int) -> np.ndarray:
    """Tile-row boundaries of `world` equal blocks (the blocks the weights are computed on, in parallel)."""
    return np.asarray([tilem * k // world for k in range(world + 1)], dtype=np.int64)


def exchange_weights(local_w: np.ndarray, blocks: np.ndarray, dist=None, device="cuda") -> np.ndarray:
    """All-gather the per-block weight vectors (block k = tile-rows [blocks[k], blocks[k+1]) computed by rank k) into the
    full weight vector. Blocks may differ in length by one: padded to the longest."""
    import torch
    rank, world = _world(dist)
    if world == 1:
        return np.asarray(local_w, dtype=np.int64)
    longest = int(np.max(np.diff(blocks)))
    t = torch.zeros(max(longest, 1), dtype=torch.int64, device=_comm_device(dist, device))
    lw = np.asarray(local_w, dtype=np.int64)
    if lw.size:
        t[:lw.size] = torch.from_numpy(lw).to(t.device)
    outs = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return np.concatenate([o.cpu().numpy()[:int(blocks[k + 1] - blocks[k])] for k, o in enumerate(outs)])


def csr_row_slice(rowptr: np.ndarray, colidx: np.ndarray, val: np.ndarray, r0: int, r1: int):
    """Rows [r0, r1) of a host CSR matrix as a CSR of its own (row pointer rebased). The device path does not use it
    (DeviceCSR.row_slice); kept for host-side tests."""
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


def _al(x, a=256):
    return (x + a - 1) // a * a


@dataclass
class Shard:
    """What one rank holds after distribute(): the whole tiled B, its own tile-rows of A, and the partition."""
    rank: int
    world: int
    m: int
    n: int
    nB: int
    aat: bool
    nnzA: int
    cuts: np.ndarray                 # tile-row boundaries of the ranks, len world + 1
    imbalance: float
    dA_full: object                  # DeviceCSR, whole A (wraps the broadcast buffer)
    dB: object                       # DeviceCSR of B (dA_full itself for C = A^2)
    tB: object                       # DeviceTiled, whole B, col-major
    dA: object                       # DeviceCSR, this rank's rows of A (borrowed slice of dA_full)
    tA: object                       # DeviceTiled, this rank's tile-rows of A
    nnzCub: int
    bcast_ms: float = 0.0
    bcast_bytes: int = 0
    mode: str = "csr"
    keep: list = field(default_factory=list)   # torch tensors whose memory the library borrows

    @property
    def trow0(self) -> int:
        return int(self.cuts[self.rank])

    @property
    def trow1(self) -> int:
        return int(self.cuts[self.rank + 1])

    def free(self):
        for o in (self.tA, self.dA, self.tB):
            o.free()
        if self.dB is not self.dA_full:
            self.dB.free()
        self.dA_full.free()
        self.keep.clear()


def distribute(A_host, aat: bool, dist=None, device=None, mode: str = "csr") -> Shard:
    """Collective. A_host = (m, n, rowptr, colidx, val) numpy arrays on rank 0 (ignored elsewhere). Returns this rank's
    Shard. `device` is this rank's torch.device; `dist` the initialised torch.distributed module (None: single GPU)."""
    import torch
    from . import api

    rank, world = _world(dist)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())

    def bcast(t):
        if world > 1:
            dist.broadcast(t, 0)
        return t

    # sizes
    hdr = torch.zeros(3, dtype=torch.int64, device=device)
    if rank == 0:
        m, n, rp, ci, v = A_host
        hdr = torch.tensor([int(m), int(n), int(rp[m])], dtype=torch.int64, device=device)
    m, n, nnz = (int(x) for x in bcast(hdr).cpu())
    # A's CSR in ONE buffer: rowptr | colidx | val, each 256-byte aligned
    o_ci = _al((m + 1) * 4)
    o_v = o_ci + _al(max(nnz, 1) * 4)
    nbytes = o_v + max(nnz, 1) * 8
    buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
    t_rp = buf[:(m + 1) * 4].view(torch.int32)
    t_ci = buf[o_ci:o_ci + max(nnz, 1) * 4].view(torch.int32)
    t_v = buf[o_v:o_v + max(nnz, 1) * 8].view(torch.float64)
    if rank == 0:
        t_rp.copy_(torch.from_numpy(np.ascontiguousarray(rp, np.int32)))
        if nnz:
            t_ci[:nnz].copy_(torch.from_numpy(np.ascontiguousarray(ci, np.int32)))
            t_v[:nnz].copy_(torch.from_numpy(np.ascontiguousarray(v, np.float64)))
    torch.cuda.synchronize(device)
    bcast_ms, bcast_bytes = 0.0, 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1 and mode == "csr":
        dist.barrier()
        e0.record()
        dist.broadcast(buf, 0)          # the one exchange step of the path: B (as its CSR) over NVLink
        e1.record()
        torch.cuda.synchronize(device)
        bcast_ms, bcast_bytes = e0.elapsed_time(e1), nbytes
    elif world > 1:
        dist.broadcast(buf, 0)          # mode "tiled": A still has to reach every rank; the timed exchange is B's slab below
        torch.cuda.synchronize(device)
    dA_full = api.DeviceCSR.wrap(m, n, nnz, t_rp.data_ptr(), t_ci.data_ptr(), t_v.data_ptr())
    keep = [buf]
    dB = api.transpose(dA_full) if aat else dA_full
    nB = dB.n
    if mode == "csr" or world == 1:
        tB = api.csr2tile(dB, True)
    else:  # rank 0 tiles B and broadcasts the tiled matrix as one buffer
        sizes = torch.zeros(3, dtype=torch.int64, device=device)
        if rank == 0:
            tB = api.csr2tile(dB, True)
            sizes = torch.tensor([tB.numtile, tB.nnz, tB.d.npat], dtype=torch.int64, device=device)
        nt, bnnz, npat = (int(x) for x in bcast(sizes).cpu())
        if rank != 0:
            tB = api.tile_alloc(dB.m, dB.n, nt, bnnz, True)
            tB.d.npat = npat  # the pattern ids travel inside the slab; their count does not
        slab = tile_slab_tensor(tB, device)
        torch.cuda.synchronize(device)
        dist.barrier()
        e0.record()
        dist.broadcast(slab, 0)
        e1.record()
        torch.cuda.synchronize(device)
        bcast_ms, bcast_bytes = e0.elapsed_time(e1), int(slab.numel())
    # step-1 weights, in parallel over equal blocks of tile-rows
    tilem = (m + 15) // 16
    blocks = equal_row_blocks(tilem, world)
    b0, b1 = int(blocks[rank]) * 16, min(int(blocks[rank + 1]) * 16, m)
    if b1 > b0:
        blk = dA_full.row_slice(b0, b1)
        tblk = api.csr2tile(blk, False)
        lw = api.tilerow_weights(tblk, tB)
        tblk.free()
        blk.free()
    else:
        lw = np.zeros(0, np.int64)
    w = exchange_weights(lw, blocks, dist, device)
    cuts = partition_tilerows(w, world)
    r0, r1 = int(cuts[rank]) * 16, min(int(cuts[rank + 1]) * 16, m)
    dA = dA_full.row_slice(r0, max(r1, r0))
    tA = api.csr2tile(dA, False)
    nnzCub = api.nnzcub(dA_full, dB)
    return Shard(rank=rank, world=world, m=m, n=n, nB=nB, aat=bool(aat), nnzA=nnz, cuts=cuts, imbalance=imbalance(w, cuts),
                 dA_full=dA_full, dB=dB, tB=tB, dA=dA, tA=tA, nnzCub=int(nnzCub), bcast_ms=bcast_ms, bcast_bytes=bcast_bytes,
                 mode=mode, keep=keep)


def spgemm(shard: Shard, slab_pairs: int | None = None, sink=None, weights=None):
    """Steps 1-3 over this rank's C tile-rows (no communication). One slab unless `slab_pairs` bounds the tile pairs per
    slab (R-MAT: C does not fit int32 offsets / one GPU whole). `sink(C_slab, stats)` sees every slab before it is freed;
    C_slab.trow0 is relative to this rank's first tile-row (add shard.trow0 for the global one).
    Returns (totals, per-slab stats) like api.spgemm_slabs."""
    from . import api
    if not slab_pairs:
        c, st = api.spgemm(shard.tA, shard.tB)
        if sink is not None:
            sink(c, st)
        c.free()
        return dict(st, slabs=1), [dict(st, trow0=0, trow1=shard.tA.tilem)]
    return api.spgemm_slabs(shard.tA, shard.tB, max_pairs=slab_pairs, sink=sink, weights=weights)


def concat(shard: Shard, totals: dict, dist=None, device="cuda") -> dict:
    """"Concatenation" of the distributed C: all-gather the per-rank (tiles, nnz, rows) and return this rank's 64-bit
    offsets inside the whole C plus the global totals (tile_ptr / tile_nnz / rowptr of a rank are rebased by them)."""
    rows = min(shard.trow1 * 16, shard.m) - min(shard.trow0 * 16, shard.m)
    counts = gather_counts([totals["numblkC"], totals["nnzC"], rows], dist, device)
    offs = concat_offsets(counts)
    return {"tile_offset": int(offs[shard.rank, 0]), "nnz_offset": int(offs[shard.rank, 1]), "row_offset": int(offs[shard.rank, 2]),
            "numblkC": int(counts[:, 0].sum()), "nnzC": int(counts[:, 1].sum()), "per_rank": counts}


def gather_csr(shard: Shard, local_csr, dist=None, device=None):
    """Whole CSR(C) on rank 0 from every rank's local CSR (DeviceCSR of its rows, row pointer starting at 0): grouped
    point-to-point transfers over NCCL, row pointers rebased in 64 bits. Returns (rowptr int64, colidx, val) numpy arrays
    on rank 0, None elsewhere. For verification and small results: C normally stays distributed."""
    import torch
    rank, world = _world(dist)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    counts = gather_counts([local_csr.m, local_csr.nnz], dist, device)
    if world == 1:
        rp, ci, v = local_csr.download()
        return rp.astype(np.int64), ci, v
    d = local_csr.d
    nz = max(int(local_csr.nnz), 1)
    mine = [torch.as_tensor(DeviceBuffer(d.rowptr, (local_csr.m + 1) * 4), device=device).view(torch.int32),
            torch.as_tensor(DeviceBuffer(d.colidx, nz * 4), device=device).view(torch.int32),
            torch.as_tensor(DeviceBuffer(d.val, nz * 8), device=device).view(torch.float64)]
    if rank != 0:
        ops = [dist.P2POp(dist.isend, t, 0) for t in mine]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        torch.cuda.synchronize(device)
        return None
    parts = [[t.clone() for t in mine]]
    ops = []
    for src in range(1, world):
        rows, nnz = int(counts[src, 0]), max(int(counts[src, 1]), 1)
        bufs = [torch.empty(rows + 1, dtype=torch.int32, device=device), torch.empty(nnz, dtype=torch.int32, device=device),
                torch.empty(nnz, dtype=torch.float64, device=device)]
        parts.append(bufs)
        ops += [dist.P2POp(dist.irecv, t, src) for t in bufs]
    for r in dist.batch_isend_irecv(ops):
        r.wait()
    torch.cuda.synchronize(device)
    offs = concat_offsets(counts)
    rowptr = np.zeros(int(counts[:, 0].sum()) + 1, np.int64)
    cols, vals = [], []
    for k, (rp, ci, v) in enumerate(parts):
        rows, nnz = int(counts[k, 0]), int(counts[k, 1])
        r0 = int(offs[k, 0])
        rowptr[r0:r0 + rows + 1] = rp.cpu().numpy().astype(np.int64) + int(offs[k, 1])
        cols.append(ci.cpu().numpy()[:nnz])
        vals.append(v.cpu().numpy()[:nnz])
    return rowptr, np.concatenate(cols), np.concatenate(vals)


def stitch_slab_csr(pieces):
    """One CSR out of the CSRs of consecutive row slabs (each (rowptr, colidx, val) with rowptr[0] == 0), row pointers
    rebased in 64 bits. Pure numpy (covered on CPU)."""
    if not pieces:
        return np.zeros(1, np.int64), np.zeros(0, np.int32), np.zeros(0, np.float64)
    offs = np.cumsum([0] + [int(len(p[1])) for p in pieces[:-1]], dtype=np.int64)
    rp = np.concatenate([np.zeros(1, np.int64)] + [np.asarray(p[0][1:], np.int64) + o for p, o in zip(pieces, offs)])
    return rp, np.concatenate([p[1] for p in pieces]), np.concatenate([p[2] for p in pieces])


def run(A_host, aat: bool, dist=None, device=None, mode: str = "csr", slab_pairs: int | None = None, gather: bool = False) -> dict:
    """Collective: the whole multi-GPU product in one call -- SURVEY.md 8(b)'s additive `tilespgemm_multi` (the
    reference picks one device, src/main.cu:41-49). A_host = (m, n, rowptr, colidx, val) on rank 0; C = A^2 or, with
    aat, A A^T. Every rank computes its tile-rows of C (distribute -> steps 1-3 -> tile2csr, slab by slab when
    `slab_pairs` bounds the tile pairs of a slab) and keeps them as host CSR: the returned dict holds
      "local_csr": (rowptr int64, colidx, val) of this rank's rows, "row0"/"row1": which rows of C they are,
      "offsets": concat()'s 64-bit offsets of this rank inside the whole C and the global totals,
      "csr": the whole CSR(C) on rank 0 when `gather` (None elsewhere and otherwise), "nnzCub", "cuts", "imbalance".
    C larger than 2^31-1 nonzeros per rank cannot be gathered into one int32-indexed device CSR: leave `gather` off."""
    from . import api
    sh = distribute(A_host, aat, dist, device, mode=mode)
    try:
        pieces = []

        def sink(tC, st):
            csr = api.tile2csr_device(tC)
            pieces.append(csr.download())
            csr.free()

        totals, _ = spgemm(sh, slab_pairs, sink)
        rp, ci, v = stitch_slab_csr(pieces)
        row0, row1 = min(sh.trow0 * 16, sh.m), min(sh.trow1 * 16, sh.m)
        if len(rp) != row1 - row0 + 1 or int(rp[-1]) != int(totals["nnzC"]):
            raise RuntimeError(f"multigpu.run: rank {sh.rank} assembled {len(rp) - 1} rows / {int(rp[-1])} nonzeros, "
                               f"expected {row1 - row0} / {int(totals['nnzC'])}")
        offsets = concat(sh, totals, dist, device if device is not None else "cuda")
        whole = None
        if gather:
            if int(rp[-1]) >= 2 ** 31:
                raise OverflowError("multigpu.run(gather=True): this rank's part of C passes 2^31-1 nonzeros")
            local = api.DeviceCSR.upload(row1 - row0, sh.nB, rp.astype(np.int32), ci, v)
            try:
                whole = gather_csr(sh, local, dist, device)
            finally:
                local.free()
        return {"local_csr": (rp, ci, v), "row0": row0, "row1": row1, "offsets": offsets, "csr": whole, "nnzCub": sh.nnzCub,
                "cuts": sh.cuts.copy(), "imbalance": sh.imbalance, "bcast_ms": sh.bcast_ms, "bcast_bytes": sh.bcast_bytes}
    finally:
        sh.free()
