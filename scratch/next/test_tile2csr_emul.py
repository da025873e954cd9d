"""Host emulation of tile2csr_v2.cuh against the oracle's CSR(C): whole matrices and slabs of tile-rows, empty tiles included.
usage: make -C scratch/next && python scratch/next/test_tile2csr_emul.py"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import oracle as orc  # noqa: E402
from spgemm_b200 import matrices as M  # noqa: E402

lib = C.CDLL(os.path.join(HERE, "libtile2csr_emul.so"))


def p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def run(name, m, n, A, t0=0, t1=None, base=0):
    tA, tB = orc.csr2tile_row_major(m, n, *A), orc.csr2tile_col_major(m, n, *A)
    t1 = tA.tilem if t1 is None else t1
    r0, r1 = t0 * 16, min(t1 * 16, m)
    sub = orc.spgemm_spa(A, A, n, r0, r1)
    tC = orc.ctiles_from_csr(m, n, tA, tB, sub, t0, t1)
    rows = r1 - r0
    arr = lambda x, dt: np.ascontiguousarray(x, dt)  # noqa: E731
    nz = max(int(tC.nnz), 1)
    rowptr, oc, ov = np.full(rows + 1, -7, np.int32), np.full(nz, -7, np.int32), np.full(nz, -7.0)
    tr = arr(tC.tile_rowidx, np.int32)
    assert tC.numtile == 0 or (tr.min() >= t0 and tr.max() < t1), "tile_rowidx is absolute"
    lib.emul_tile2csr(rows, t1 - t0, t0, tC.numtile, p(arr(tC.tile_ptr, np.int32), C.c_int), p(tr if tC.numtile else np.zeros(1, np.int32), C.c_int),
                      p(arr(tC.tile_columnidx, np.int32) if tC.numtile else np.zeros(1, np.int32), C.c_int), p(arr(tC.tile_nnz, np.int32), C.c_int),
                      p(arr(tC.ptr, np.uint16) if tC.numtile else np.zeros(1, np.uint16), C.c_uint16),
                      p(arr(tC.col, np.uint16) if tC.nnz else np.zeros(1, np.uint16), C.c_uint16),
                      p(arr(tC.val, np.float64) if tC.nnz else np.zeros(1), C.c_double), base, p(rowptr, C.c_int), p(oc, C.c_int), p(ov, C.c_double))
    assert np.array_equal(rowptr, sub[0] + base), name + " rowptr"
    assert np.array_equal(oc[:tC.nnz], sub[1]) and np.array_equal(ov[:tC.nnz], sub[2]), name + " col/val"
    print(f"{name:34s} ok: rows {rows}, C tiles {tC.numtile} ({int((np.diff(tC.tile_nnz) == 0).sum())} empty), nnz {tC.nnz}")


CASES = {
    "lap2d_48": lambda: M.lap2d(48), "lap2d_33x17": lambda: M.lap2d(33, 17), "stencil27_9": lambda: M.stencil27(9),
    "stencil27_20x7x5": lambda: M.stencil27(20, 7, 5), "blockfem_120": lambda: M.blockfem(120), "rmat_s10": lambda: M.rmat(10, 8, seed=5),
    "rmat_s12_hypersparse": lambda: M.rmat(12, 2, a=.3, b=.25, c=.25, d=.2, seed=2),
    "rand_ragged_203": lambda: M.random_sparse(203, 203, 0.03, seed=11), "full_48": lambda: M.random_sparse(48, 48, 5.0, seed=15),
    "empty": lambda: (33, 33, np.zeros(34, np.int32), np.zeros(0, np.int32), np.zeros(0)),
    "one_by_one": lambda: (1, 1, np.array([0, 1], np.int32), np.array([0], np.int32), np.array([2.0])),
}
for name, gen in CASES.items():
    m, n, rp, ci, _ = gen()
    v = M.set_values(len(ci), "hash") if len(ci) else np.zeros(0)
    run(name, m, n, (rp, ci, v))
m, n, rp, ci, _ = M.stencil27(12)
v = M.set_values(len(ci), "hash")
tilem = (m + 15) // 16
for t0, t1 in ((0, 5), (5, 6), (6, 40), (40, tilem)):
    run(f"stencil27_12 slab [{t0},{t1}) base 1000", m, n, (rp, ci, v), t0, t1, base=1000)
print("all tile2csr-emulation cases passed")
