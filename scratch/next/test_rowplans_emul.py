"""Host emulation of rowplans.cuh (step 1 from tile-row recipes) against the tile-level product computed with numpy:
C tile list (columns, tile-rows, c_tile_ptr), pair ranges and the (A tile, B storage id) pairs in ascending-K order -- exactly
what k_step1 emits today. Whole matrices and slabs.   usage: make -C scratch/next && python scratch/next/test_rowplans_emul.py"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import oracle as orc  # noqa: E402
from spgemm_b200 import matrices as M  # noqa: E402

lib = C.CDLL(os.path.join(HERE, "librowplans_emul.so"))


def p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def expected(tA, tB, t0, t1):
    a_ptr, a_col = tA.tile_ptr.astype(np.int64), tA.tile_columnidx.astype(np.int64)
    b_ptr, b_col = tB.tile_ptr.astype(np.int64), tB.tile_columnidx.astype(np.int64)
    b_row = np.repeat(np.arange(tB.tilem, dtype=np.int64), np.diff(b_ptr))
    order = np.lexsort((b_row, b_col))
    rm2csc = np.empty(len(b_col), np.int64)
    rm2csc[order] = np.arange(len(b_col))
    a_row = np.repeat(np.arange(tA.tilem, dtype=np.int64), np.diff(a_ptr))
    sel = np.flatnonzero((a_row >= t0) & (a_row < t1))
    cnt = b_ptr[a_col[sel] + 1] - b_ptr[a_col[sel]]
    src = np.repeat(sel, cnt)
    dst = np.repeat(b_ptr[a_col[sel]], cnt) + (np.arange(int(cnt.sum())) - np.repeat(np.cumsum(cnt) - cnt, cnt))
    ckey = a_row[src] * tB.tilen + b_col[dst]
    o = np.lexsort((a_col[src], ckey))
    ckey, pa, pb = ckey[o], src[o], rm2csc[dst[o]]
    first = np.flatnonzero(np.r_[True, ckey[1:] != ckey[:-1]]) if len(ckey) else np.zeros(0, np.int64)
    crow, ccol = ckey[first] // tB.tilen, ckey[first] % tB.tilen
    cptr = np.searchsorted(crow, np.arange(t0, t1 + 1))
    return rm2csc.astype(np.int32), cptr, ccol, crow, np.r_[first, len(ckey)], pa, pb


def run(name, m, n, A, B, nB, t0=0, t1=None, must_fail=False):
    tA = orc.csr2tile_row_major(m, n, *A)
    tB = orc.csr2tile_col_major(len(B[0]) - 1, nB, *B)
    t1 = tA.tilem if t1 is None else t1
    rm2csc, cptr, ccol, crow, pptr, pa, pb = expected(tA, tB, t0, t1)
    nC, nP, ntr = len(ccol), len(pa), t1 - t0
    arr = lambda x: np.ascontiguousarray(x, np.int32) if len(x) else np.zeros(1, np.int32)  # noqa: E731
    o_cptr = np.full(ntr + 1, -7, np.int32)
    o = [np.full(max(nC, 1), -7, np.int32) for _ in range(4)] + [np.full(max(nP, 1), -7, np.int32) for _ in range(2)]
    info = np.zeros(4, np.int64)
    rc = lib.emul_rowplans(tA.tilem, p(arr(tA.tile_ptr), C.c_int), p(arr(tA.tile_columnidx), C.c_int), tB.tilem, p(arr(tB.tile_ptr), C.c_int),
                           p(arr(tB.tile_columnidx), C.c_int), p(arr(rm2csc), C.c_int), t0, ntr, p(o_cptr, C.c_int), p(o[0], C.c_int),
                           p(o[1], C.c_int), p(o[2], C.c_int), p(o[3], C.c_int), p(o[4], C.c_int), p(o[5], C.c_int),
                           C.c_longlong(max(nC, 1)), C.c_longlong(max(nP, 1)), p(info, C.c_longlong))
    if must_fail:
        assert rc == 1, (name, rc)
        print(f"{name:36s} fell back as it must (tile-rows with > 4096 pairs or too many recipes)")
        return
    assert rc == 0, (name, rc)
    assert np.array_equal(o_cptr, cptr), name + " c_tile_ptr"
    assert np.array_equal(o[0][:nC], ccol) and np.array_equal(o[1][:nC], crow), name + " C tile list"
    assert np.array_equal(o[2][:nC], pptr[:-1]) and np.array_equal(o[3][:nC], pptr[1:]), name + " pair ranges"
    assert np.array_equal(o[4][:nP], pa) and np.array_equal(o[5][:nP], pb), name + " pairs"
    print(f"{name:36s} ok: tile-rows {ntr}, B-row recipes {info[0]}, A-row recipes {info[1]}, C tiles {nC}, pairs {nP}")


CASES = {
    "lap2d_48": lambda: M.lap2d(48), "lap2d_33x17": lambda: M.lap2d(33, 17), "stencil27_9": lambda: M.stencil27(9),
    "stencil27_20x7x5": lambda: M.stencil27(20, 7, 5), "stencil27_32": lambda: M.stencil27(32), "blockfem_120": lambda: M.blockfem(120),
    "blockfem_band3": lambda: M.blockfem(40, dof=6, band=3), "rmat_s10_mild": lambda: M.rmat(10, 4, a=.3, b=.25, c=.25, d=.2, seed=5),
    "rand_ragged_203": lambda: M.random_sparse(203, 203, 0.03, seed=11), "full_48": lambda: M.random_sparse(48, 48, 5.0, seed=15),
    "empty": lambda: (33, 33, np.zeros(34, np.int32), np.zeros(0, np.int32), np.zeros(0)),
    "one_by_one": lambda: (1, 1, np.array([0, 1], np.int32), np.array([0], np.int32), np.array([2.0])),
}
for name, gen in CASES.items():
    m, n, rp, ci, v = gen()
    run(name, m, n, (rp, ci, v), (rp, ci, v), n)
m, n, rp, ci, v = M.stencil27(12)
tilem = (m + 15) // 16
for t0, t1 in ((0, 5), (5, 6), (6, 40), (40, tilem)):
    run(f"stencil27_12 slab [{t0},{t1})", m, n, (rp, ci, v), (rp, ci, v), n, t0, t1)
m, n, rp, ci, v = M.rmat(10, 8, seed=9)
cp, ri, cv = orc.transpose(m, n, rp, ci, v)
run("rmat_s10 AA^T", m, n, (rp, ci, v), (cp, ri, cv), m)
m, n, rp, ci, v = M.rmat(13, 16, seed=1)
run("rmat_s13 skewed (hub tile-rows)", m, n, (rp, ci, v), (rp, ci, v), n, must_fail=True)
m, k, rpA, ciA, vA = M.random_sparse(70, 100, 0.05, seed=21)
_, n2, rpB, ciB, vB = M.random_sparse(100, 45, 0.06, seed=22)
run("rectangular 70x100x45", m, k, (rpA, ciA, vA), (rpB, ciB, vB), n2)
print("all row-plan emulation cases passed")
if len(sys.argv) > 1 and sys.argv[1] == "big":     # statistics on larger structured cases (tens of seconds)
    for name, gen in {"stencil27_64": lambda: M.stencil27(64), "stencil27_100": lambda: M.stencil27(100), "lap2d_1000": lambda: M.lap2d(1000),
                      "blockfem_100000": lambda: M.blockfem(100000)}.items():
        m, n, rp, ci, v = gen()
        run(name, m, n, (rp, ci, v), (rp, ci, v), n)
