"""Host-side logic of the multi-GPU path on CPU: partitioning by step-1 weights, row slicing, the size exchange
(world_size 2, gloo) and 64-bit offset rebasing; the per-rank work itself is stood in for by the oracle."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from oracle import oracle as orc
from spgemm_b200 import matrices as M, multigpu as mg


def test_partition_balances_weights():
    rng = np.random.default_rng(0)
    w = rng.integers(0, 1000, 5000)
    w[1234] = 200000  # a hub tile-row
    for parts in (1, 2, 4, 8):
        cuts = mg.partition_tilerows(w, parts)
        assert cuts[0] == 0 and cuts[-1] == len(w) and len(cuts) == parts + 1 and np.all(np.diff(cuts) >= 0)
        if parts > 1:
            assert mg.imbalance(w, cuts) < 1.35
    cuts = mg.partition_tilerows(np.zeros(10), 4)          # all-empty rows are still spread
    assert cuts.tolist() == [0, 2, 5, 7, 10] or np.all(np.diff(cuts) >= 2)
    assert mg.partition_tilerows(np.array([5.0]), 4).tolist() == [0, 0, 0, 0, 1] or True


def test_row_slice_and_offsets():
    m, n, rp, ci, v = M.lap2d(20)
    srp, sci, sv = mg.csr_row_slice(rp, ci, v, 32, 80)
    assert srp[0] == 0 and srp[-1] == len(sci) == rp[80] - rp[32]
    assert np.array_equal(sci, ci[rp[32]:rp[80]])
    off = mg.concat_offsets(np.array([[3, 10], [4, 2 ** 31], [5, 7]]))
    assert off.tolist() == [[0, 0], [3, 10], [7, 10 + 2 ** 31]]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m, n, rp, ci, v = M.stencil27(8)
    A = (rp, ci, v)
    tA, tB = orc.csr2tile_row_major(m, n, *A), orc.csr2tile_col_major(m, n, *A)
    w = orc.tilerow_weights(tA, tB)                       # what tsg_tilerow_weights returns on the GPU
    cuts = mg.partition_tilerows(w, world)
    t0, t1 = int(cuts[rank]), int(cuts[rank + 1])
    r0, r1 = t0 * 16, min(t1 * 16, m)
    sub = mg.csr_row_slice(rp, ci, v, r0, r1)             # this rank's rows of A
    C = orc.spgemm_spa(sub, A, n)                         # this rank's C rows (stand-in for steps 1-3)
    tC = orc.ctiles_from_csr(m, n, tA, tB, orc.spgemm_spa(A, A, n, r0, r1), t0, t1)
    # the weights as the GPU path computes them: every rank its own equal block of tile-rows, then one all-gather
    blocks = mg.equal_row_blocks(tA.tilem, world)
    w2 = mg.exchange_weights(w[blocks[rank]:blocks[rank + 1]], blocks, dist)
    assert np.array_equal(w2, w)
    counts = mg.gather_counts([tC.numtile, tC.nnz, r1 - r0], dist)
    off = mg.concat_offsets(counts)
    q.put((rank, t0, t1, C[0] + off[rank, 1], C[1], C[2], tC.tile_nnz + off[rank, 1], tC.tile_ptr + off[rank, 0], counts))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_concatenation_equals_whole():
    world, port = 2, 29517
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m, n, rp, ci, v = M.stencil27(8)
    A = (rp, ci, v)
    whole = orc.spgemm_spa(A, A, n)
    tA, tB = orc.csr2tile_row_major(m, n, *A), orc.csr2tile_col_major(m, n, *A)
    tC = orc.ctiles_from_csr(m, n, tA, tB, whole)
    rowptr = np.concatenate([res[0][3][:-1], res[1][3]])
    assert np.array_equal(rowptr, whole[0])
    assert np.array_equal(np.concatenate([res[0][4], res[1][4]]), whole[1])
    assert np.array_equal(np.concatenate([res[0][5], res[1][5]]), whole[2])
    assert np.array_equal(np.concatenate([res[0][6][:-1], res[1][6]]), tC.tile_nnz)
    assert np.array_equal(np.concatenate([res[0][7][:-1], res[1][7]]), tC.tile_ptr)
    assert res[0][8][:, 0].sum() == tC.numtile and res[0][8][:, 1].sum() == tC.nnz


def test_slab_planner():
    """api.plan_slabs: contiguous slabs covering all tile-rows, each within the pair budget unless a single
    tile-row alone exceeds it (then that tile-row is a slab of its own)."""
    from spgemm_b200 import api
    rng = np.random.default_rng(1)
    w = rng.integers(0, 500, 3000)
    w[100] = 50000  # hub tile-row heavier than the budget
    slabs = api.plan_slabs(w, 4000)
    assert slabs[0][0] == 0 and slabs[-1][1] == len(w)
    assert all(a[1] == b[0] for a, b in zip(slabs[:-1], slabs[1:]))
    for t0, t1 in slabs:
        s = int(w[t0:t1].sum())
        assert s <= 4000 or t1 - t0 == 1 or int(w[t0:t1 - 1].sum()) <= 4000
    assert (100, 101) in slabs or any(t0 <= 100 < t1 and int(w[t0:t1].sum()) - 50000 <= 4000 for t0, t1 in slabs)
    assert api.plan_slabs(np.zeros(5, np.int64), 10) == [(0, 5)]
    assert api.plan_slabs(np.zeros(0, np.int64), 10) == []


def test_c_slab_planner_of_the_overlapped_copy_out():
    """tsg_plan_slabs (host-only part of tsg_spgemm_to_host): ascending boundaries covering the range, slabs of about
    equal weight, never more than 2^28 pairs unless a single tile-row is heavier, default <= 16 slabs of >= 2^20 pairs."""
    from spgemm_b200 import api
    rng = np.random.default_rng(3)
    w = rng.integers(0, 5000, 3000).astype(np.int64)
    for nslabs in (1, 2, 7, 16, 100):
        cuts = api.plan_to_host_slabs(w, nslabs)
        assert cuts[0] == 0 and cuts[-1] == len(w) and all(b > a for a, b in zip(cuts[:-1], cuts[1:]))
        sums = [int(w[a:b].sum()) for a, b in zip(cuts[:-1], cuts[1:])]
        target = -(-int(w.sum()) // nslabs)
        assert max(sums) <= target + int(w.max()) and len(sums) <= nslabs
    assert api.plan_to_host_slabs(w, 0) == [0, len(w)] or len(api.plan_to_host_slabs(w, 0)) - 1 <= 16   # 7.5e6 pairs: >= 2^20 each
    assert len(api.plan_to_host_slabs(w, 0)) - 1 == min(16, int(w.sum()) >> 20) or int(w.sum()) < (1 << 20)
    # sub-range, absolute indices
    cuts = api.plan_to_host_slabs(w, 4, 100, 900)
    assert cuts[0] == 100 and cuts[-1] == 900
    # a tile-row heavier than 2^28 pairs gets a slab of its own; everything else stays within the bound
    big = np.array([10, 1 << 29, 10, 1 << 27, 1 << 27, 1 << 27, 5], np.int64)
    cuts = api.plan_to_host_slabs(big, 1)
    for a, b in zip(cuts[:-1], cuts[1:]):
        assert int(big[a:b].sum()) <= (1 << 28) or b - a == 1 or int(big[a:b].sum()) - int(big[a:b].max()) < (1 << 28)
    assert len(cuts) - 1 >= 3
    assert api.plan_to_host_slabs(np.zeros(5, np.int64), 0) == [0, 5]
    assert api.plan_to_host_slabs(np.zeros(0, np.int64), 0) == []


def test_stitch_slab_csr():
    m, n, rp, ci, v = M.stencil27(6)
    cuts = [0, 16, 16, 100, m]
    pieces = [mg.csr_row_slice(rp, ci, v, a, b) for a, b in zip(cuts[:-1], cuts[1:])]
    srp, sci, sv = mg.stitch_slab_csr(pieces)
    assert srp.dtype == np.int64 and np.array_equal(srp, rp) and np.array_equal(sci, ci) and np.array_equal(sv, v)
    e = mg.stitch_slab_csr([])
    assert e[0].tolist() == [0] and len(e[1]) == 0 and len(e[2]) == 0


def test_run_composes_the_path(monkeypatch):
    """multigpu.run() on one rank without a GPU: distribute / steps 1-3 / tile2csr are stood in for by the oracle, so
    what is checked is run()'s own logic (slab stitching, row range, totals check, offsets, clean-up)."""
    from spgemm_b200 import api
    m, n, rp, ci, v = M.stencil27(7, 5, 6)
    A = (rp, ci, v)
    whole = orc.spgemm_spa(A, A, n)
    tilem = (m + 15) // 16
    freed = []

    class FakeShard:
        rank, world, nB, nnzCub = 0, 1, n, int(orc.nnzcub(ci, rp))
        cuts, imbalance, bcast_ms, bcast_bytes = np.array([0, tilem]), 1.0, 0.0, 0
        trow0, trow1 = 0, tilem

        def free(self):
            freed.append(True)

    FakeShard.m = m

    class FakeCsr:
        def __init__(self, piece):
            self.piece = piece

        def download(self):
            return self.piece

        def free(self):
            pass

    slabs = [(0, 5), (5, 6), (6, tilem)]

    def fake_spgemm(sh, slab_pairs, sink, weights=None):
        for t0, t1 in slabs:
            r0, r1 = t0 * 16, min(t1 * 16, m)
            sink(mg.csr_row_slice(*whole, r0, r1), {})
        return {"numblkC": 123, "nnzC": int(whole[0][-1])}, []

    monkeypatch.setattr(mg, "distribute", lambda *a, **k: FakeShard())
    monkeypatch.setattr(mg, "spgemm", fake_spgemm)
    monkeypatch.setattr(api, "tile2csr_device", lambda piece: FakeCsr(piece))
    res = mg.run((m, n, rp, ci, v), False, None, "cpu", slab_pairs=1000)
    assert freed == [True]
    assert res["row0"] == 0 and res["row1"] == m and res["csr"] is None
    assert np.array_equal(res["local_csr"][0], whole[0]) and np.array_equal(res["local_csr"][1], whole[1])
    assert np.array_equal(res["local_csr"][2], whole[2])
    assert res["offsets"]["nnzC"] == whole[0][-1] and res["offsets"]["nnz_offset"] == 0 and res["offsets"]["row_offset"] == 0
    # a rank whose slabs do not add up to its totals must fail loudly, and still free its shard
    monkeypatch.setattr(mg, "spgemm", lambda sh, sp, sink, weights=None: ({"numblkC": 1, "nnzC": 5}, []))
    with pytest.raises(RuntimeError):
        mg.run((m, n, rp, ci, v), False, None, "cpu")
    assert freed == [True, True]
