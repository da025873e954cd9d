"""The three drafts chained under host emulation, the way round 2 would run them on the device:
rowplans (step 1: C tile list + pair lists) -> plans (symbolic + numeric) -> tile2csr_v2 -> CSR(C), compared with the
oracle's serial SPA. Interfaces between the drafts are the product's own arrays (tile list, pair_ptr/pair_end/pair_a/pair_b,
mask/Ptr/tile_nnz/Col/Val).   usage: make -C scratch/next && python scratch/next/test_chain_emul.py"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import oracle as orc  # noqa: E402
from spgemm_b200 import matrices as M  # noqa: E402

L1 = C.CDLL(os.path.join(HERE, "librowplans_emul.so"))
L2 = C.CDLL(os.path.join(HERE, "libplans_emul.so"))
L3 = C.CDLL(os.path.join(HERE, "libtile2csr_emul.so"))
I32 = lambda x: np.ascontiguousarray(x, np.int32) if len(x) else np.zeros(1, np.int32)  # noqa: E731
U16 = lambda x: np.ascontiguousarray(x, np.uint16) if len(x) else np.zeros(1, np.uint16)  # noqa: E731
F64 = lambda x: np.ascontiguousarray(x, np.float64) if len(x) else np.zeros(1)  # noqa: E731


def p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def chain(name, m, n, A, exact):
    tA, tB = orc.csr2tile_row_major(m, n, *A), orc.csr2tile_col_major(m, n, *A)
    exp = orc.spgemm_spa(A, A, n)
    b_ptr, b_col = tB.tile_ptr.astype(np.int64), tB.tile_columnidx.astype(np.int64)
    b_row = np.repeat(np.arange(tB.tilem, dtype=np.int64), np.diff(b_ptr))
    rm2csc = np.empty(len(b_col), np.int64)
    rm2csc[np.lexsort((b_row, b_col))] = np.arange(len(b_col))
    ntr = tA.tilem
    # ---- step 1 from row recipes
    capC, capP = 1 << 22, 1 << 24
    cptr = np.zeros(ntr + 1, np.int32)
    ccol, crow, pptr, pend = (np.zeros(capC, np.int32) for _ in range(4))
    pa, pb = np.zeros(capP, np.int32), np.zeros(capP, np.int32)
    info = np.zeros(4, np.int64)
    rc = L1.emul_rowplans(tA.tilem, p(I32(tA.tile_ptr), C.c_int), p(I32(tA.tile_columnidx), C.c_int), tB.tilem, p(I32(tB.tile_ptr), C.c_int),
                          p(I32(tB.tile_columnidx), C.c_int), p(I32(rm2csc), C.c_int), 0, ntr, p(cptr, C.c_int), p(ccol, C.c_int),
                          p(crow, C.c_int), p(pptr, C.c_int), p(pend, C.c_int), p(pa, C.c_int), p(pb, C.c_int), C.c_longlong(capC),
                          C.c_longlong(capP), p(info, C.c_longlong))
    assert rc == 0, (name, "rowplans", rc)
    nC, nrow_rec = int(info[2]), int(info[1])
    # ---- symbolic + numeric from C-tile recipes
    c_mask, c_ptr, c_tn = np.zeros(max(nC, 1) * 16, np.uint16), np.zeros(max(nC, 1) * 16, np.uint16), np.zeros(nC + 1, np.int32)
    cap = max(len(exp[1]), 1)
    c_col, c_val = np.zeros(cap, np.uint16), np.zeros(cap)
    info2 = np.zeros(4, np.int64)
    rc = L2.emul_plans(tA.numtile, p(U16(tA.mask), C.c_uint16), p(U16(tA.ptr), C.c_uint16), p(I32(tA.tile_nnz), C.c_int), p(F64(tA.val), C.c_double),
                       tB.numtile, p(U16(tB.mask), C.c_uint16), p(U16(tB.ptr), C.c_uint16), p(I32(tB.tile_nnz), C.c_int), p(F64(tB.val), C.c_double),
                       nC, p(pptr, C.c_int), p(pend, C.c_int), p(pa, C.c_int), p(pb, C.c_int), p(c_mask, C.c_uint16), p(c_ptr, C.c_uint16),
                       p(c_tn, C.c_int), p(c_col, C.c_uint16), p(c_val, C.c_double), C.c_longlong(cap), p(info2, C.c_longlong))
    assert rc == 0, (name, "plans", rc)
    assert int(info2[3]) == len(exp[1]), (name, "nnzC", int(info2[3]), len(exp[1]))
    # ---- tiles -> CSR
    rowptr, oc, ov = np.zeros(m + 1, np.int32), np.zeros(cap, np.int32), np.zeros(cap)
    L3.emul_tile2csr(m, ntr, 0, nC, p(cptr, C.c_int), p(crow, C.c_int), p(ccol, C.c_int), p(c_tn, C.c_int), p(c_ptr, C.c_uint16),
                     p(c_col, C.c_uint16), p(c_val, C.c_double), 0, p(rowptr, C.c_int), p(oc, C.c_int), p(ov, C.c_double))
    nnz = len(exp[1])
    assert np.array_equal(rowptr, exp[0]) and np.array_equal(oc[:nnz], exp[1]), name + " CSR structure"
    if exact:
        assert np.array_equal(ov[:nnz], exp[2]), name + " values"
    else:
        assert np.allclose(ov[:nnz], exp[2], rtol=1e-12, atol=0), name + " values"
    print(f"{name:22s} ok: {ntr} tile-rows / {nrow_rec} row recipes, {nC} C tiles / {int(info2[1])} tile recipes, nnzC {nnz}")


for name, gen in {"lap2d_48": lambda: M.lap2d(48), "lap2d_33x17": lambda: M.lap2d(33, 17), "stencil27_9": lambda: M.stencil27(9),
                  "stencil27_20x7x5": lambda: M.stencil27(20, 7, 5), "stencil27_32": lambda: M.stencil27(32),
                  "blockfem_120": lambda: M.blockfem(120), "blockfem_band3": lambda: M.blockfem(40, dof=6, band=3),
                  "rand_ragged_203": lambda: M.random_sparse(203, 203, 0.03, seed=11)}.items():
    m, n, rp, ci, _ = gen()
    for values in ("mod10", "hash"):
        chain(f"{name}/{values}", m, n, (rp, ci, M.set_values(len(ci), values)), values == "mod10")
print("all chained emulation cases passed")
