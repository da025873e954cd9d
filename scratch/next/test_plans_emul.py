"""Host emulation of the round-2 plan kernels (plans.cuh) against the oracle: C's masks / Ptr / tile_nnz / Col bit-exact,
values bit-exact for integer-valued inputs and <= 1e-12 otherwise (same summation order as the serial SPA).
usage: make -C scratch/next && python scratch/next/test_plans_emul.py"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import oracle as orc  # noqa: E402
from spgemm_b200 import matrices as M  # noqa: E402

lib = C.CDLL(os.path.join(HERE, "libplans_emul.so"))


def p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def pair_lists(tA, tB):
    """What step 1 emits: per listed C tile (row-major order) the (A tile, B storage id) pairs in ascending K."""
    a_ptr, a_col = tA.tile_ptr.astype(np.int64), tA.tile_columnidx.astype(np.int64)
    b_ptr, b_col = tB.tile_ptr.astype(np.int64), tB.tile_columnidx.astype(np.int64)      # B's row-major tile structure
    b_row = np.repeat(np.arange(tB.tilem, dtype=np.int64), np.diff(b_ptr))
    order = np.lexsort((b_row, b_col))                                                    # CSC storage order
    rm2csc = np.empty(len(b_col), np.int64)
    rm2csc[order] = np.arange(len(b_col))
    a_row = np.repeat(np.arange(tA.tilem, dtype=np.int64), np.diff(a_ptr))
    cnt = b_ptr[a_col + 1] - b_ptr[a_col]
    src = np.repeat(np.arange(tA.numtile, dtype=np.int64), cnt)
    dst = np.repeat(b_ptr[a_col], cnt) + (np.arange(int(cnt.sum())) - np.repeat(np.cumsum(cnt) - cnt, cnt))
    ckey = a_row[src] * tB.tilen + b_col[dst]
    o = np.lexsort((a_col[src], ckey))
    ckey, pa, pb = ckey[o], src[o], rm2csc[dst[o]]
    first = np.flatnonzero(np.r_[True, ckey[1:] != ckey[:-1]]) if len(ckey) else np.zeros(0, np.int64)
    pair_ptr = np.r_[first, len(ckey)].astype(np.int32)
    return pair_ptr, pa.astype(np.int32), pb.astype(np.int32)


def run(name, m, n, A, B, nB, exact):
    tA = orc.csr2tile_row_major(m, n, *A)
    tB = orc.csr2tile_col_major(len(B[0]) - 1, nB, *B)
    exp = orc.ctiles_from_csr(m, nB, tA, tB, orc.spgemm_spa(A, B, nB))
    pair_ptr, pa, pb = pair_lists(tA, tB)
    nC = len(pair_ptr) - 1
    assert nC == exp.numtile, (name, nC, exp.numtile)
    c_mask, c_ptr = np.zeros(max(nC, 1) * 16, np.uint16), np.zeros(max(nC, 1) * 16, np.uint16)
    c_tn = np.zeros(nC + 1, np.int32)
    cap = max(int(exp.nnz), 1)
    c_col, c_val = np.zeros(cap, np.uint16), np.zeros(cap)
    info = np.zeros(4, np.int64)
    arr = lambda x, dt: np.ascontiguousarray(x, dt)  # noqa: E731
    a_mask, a_p, a_tn, a_v = arr(tA.mask, np.uint16), arr(tA.ptr, np.uint16), arr(tA.tile_nnz, np.int32), arr(tA.val, np.float64)
    b_mask, b_p, b_tn, b_v = arr(tB.mask, np.uint16), arr(tB.ptr, np.uint16), arr(tB.tile_nnz, np.int32), arr(tB.val, np.float64)
    pend = arr(pair_ptr[1:], np.int32) if nC else np.zeros(1, np.int32)
    rc = lib.emul_plans(tA.numtile, p(a_mask, C.c_uint16), p(a_p, C.c_uint16), p(a_tn, C.c_int), p(a_v, C.c_double),
                        tB.numtile, p(b_mask, C.c_uint16), p(b_p, C.c_uint16), p(b_tn, C.c_int), p(b_v, C.c_double),
                        nC, p(pair_ptr, C.c_int), p(pend, C.c_int), p(pa if len(pa) else np.zeros(1, np.int32), C.c_int),
                        p(pb if len(pb) else np.zeros(1, np.int32), C.c_int),
                        p(c_mask, C.c_uint16), p(c_ptr, C.c_uint16), p(c_tn, C.c_int), p(c_col, C.c_uint16), p(c_val, C.c_double),
                        C.c_longlong(cap), p(info, C.c_longlong))
    if rc == 1:
        print(f"{name:28s} fell back (too many patterns / recipes): A tiles {tA.numtile}, C tiles {nC}")
        return
    assert rc == 0, (name, rc)
    assert np.array_equal(c_mask[:nC * 16], exp.mask) and np.array_equal(c_ptr[:nC * 16], exp.ptr), name + " mask/Ptr"
    assert np.array_equal(c_tn, exp.tile_nnz), name + " tile_nnz"
    assert np.array_equal(c_col[:exp.nnz], exp.col), name + " Col"
    if exact:
        assert np.array_equal(c_val[:exp.nnz], exp.val), name + " values"
    else:
        assert np.allclose(c_val[:exp.nnz], exp.val, rtol=1e-12, atol=0), name + " values"
    print(f"{name:28s} ok: A tiles {tA.numtile}, C tiles {nC}, nnzC {exp.nnz}, patterns {info[0]}, recipes {info[1]}, plan entries {info[2]}")


CASES = {
    "lap2d_48": lambda: M.lap2d(48), "lap2d_33x17": lambda: M.lap2d(33, 17), "stencil27_9": lambda: M.stencil27(9),
    "stencil27_20x7x5": lambda: M.stencil27(20, 7, 5), "stencil27_32": lambda: M.stencil27(32), "blockfem_120": lambda: M.blockfem(120),
    "blockfem_band3": lambda: M.blockfem(40, dof=6, band=3), "rmat_s10": lambda: M.rmat(10, 8, seed=5),
    "rand_ragged_203": lambda: M.random_sparse(203, 203, 0.03, seed=11), "rand_dense_64": lambda: M.random_sparse(64, 64, 0.7, seed=14),
    "full_48": lambda: M.random_sparse(48, 48, 5.0, seed=15), "rmat_s13_falls_back": lambda: M.rmat(13, 16, seed=1),
    "empty": lambda: (33, 33, np.zeros(34, np.int32), np.zeros(0, np.int32), np.zeros(0)),
    "one_by_one": lambda: (1, 1, np.array([0, 1], np.int32), np.array([0], np.int32), np.array([2.0])),
}
for name, gen in CASES.items():
    m, n, rp, ci, _ = gen()
    for values in ("mod10", "hash"):
        v = M.set_values(len(ci), values) if len(ci) else np.zeros(0)
        run(f"{name}/{values}", m, n, (rp, ci, v), (rp, ci, v), n, values == "mod10")
m, n, rp, ci, v = M.rmat(10, 8, seed=9)
cp, ri, cv = orc.transpose(m, n, rp, ci, v)
run("rmat_s10 AA^T", m, n, (rp, ci, v), (cp, ri, cv), m, False)
m, k, rpA, ciA, vA = M.random_sparse(70, 100, 0.05, seed=21)
_, n2, rpB, ciB, vB = M.random_sparse(100, 45, 0.06, seed=22)
run("rectangular 70x100x45", m, k, (rpA, ciA, vA), (rpB, ciB, vB), n2, False)
print("all plan-emulation cases passed")
