"""Generate tests/golden/*.npz from the reference's OWN CPU code (oracle/_ref, compiled in place
from /root/reference/src) and from the reference's own fixtures under UnitTest/CSR2TILE.

Run here (the container that has /root/reference):   python tests/golden/make_golden.py
The vectors are committed; the GPU box has neither /root/reference nor a need for it.

Per case the file holds the input CSR and, computed by UNMODIFIED reference functions:
  A_*  csr2tile_row_major   (src/csr2tile.h:205)
  B_*  csr2tile_col_major   (src/csr2tile.h:279)
  T_*  matrix_transposition (src/utils.h:161)
  C_rowptr/C_colidx         spgemm_spa          (src/spgemm_serialref_spa_new.h:7)   [square only]
  C_val                     spgemm_serialref    (src/external/cusparse/spgemm_serialref_spa.h:33)
  C2_*                      tile2csr            (src/tile2csr.h:72) applied to the oracle's tiled C
plus, for random_0.1_36x36, the 36 golden row bitmasks of UnitTest/CSR2TILE/bitmask.h.

gtile_<case>_<tile_size_m>x<tile_size_n>.npz: the same for the fork's runtime tile sizes (src/main.cu:84-91: tiles of A are
tile_size_m x tile_size_n, tiles of B tile_size_n x tile_size_m, tiles of C tile_size_m x tile_size_m): A_* / B_* from the
unmodified csr2tile_row_major / csr2tile_col_major(matrix, tile_size_m, tile_size_n), C2_* from the unmodified
tile2csr(C, tile_size_m, tile_size_m) (src/main.cu:327) applied to the oracle's tiled C.
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc, ref  # noqa: E402
from spgemm_b200 import matrices as M  # noqa: E402

FX = "/root/reference/UnitTest/CSR2TILE/"
OUT = os.path.dirname(os.path.abspath(__file__))


def tiled_dict(prefix, t):
    d = {f"{prefix}_{k}": getattr(t, k) for k in
         ("tile_ptr", "tile_columnidx", "tile_rowidx", "tile_nnz", "val", "col", "ptr", "mask")}
    d[f"{prefix}_dims"] = np.array([t.m, t.n, t.tilem, t.tilen, t.numtile, t.nnz], dtype=np.int64)
    if t.csc_tile_ptr is not None:
        d[f"{prefix}_csc_tile_ptr"] = t.csc_tile_ptr
        d[f"{prefix}_csc_tile_rowidx"] = t.csc_tile_rowidx
    return d


def make_case(name, m, n, rp, ci, v):
    d = dict(m=np.int64(m), n=np.int64(n), rowptr=rp, colidx=ci, val=v)
    tA = ref.csr2tile_row_major(m, n, rp, ci, v)
    tB = ref.csr2tile_col_major(m, n, rp, ci, v)
    d.update(tiled_dict("A", tA))
    d.update(tiled_dict("B", tB))
    cp, ri, cv = ref.matrix_transposition(m, n, rp, ci, v)
    d.update(T_colptr=cp, T_rowidx=ri, T_val=cv)
    if m == n:
        A = (rp, ci, v)
        rpC, ciC = ref.spgemm_spa(A, A, n)
        rpC2, ciC2, vC = ref.spgemm_serialref(A, A, n)
        assert np.array_equal(rpC, rpC2) and np.array_equal(ciC, ciC2)
        d.update(C_rowptr=rpC, C_colidx=ciC, C_val=vC)
        # the reference's tile2csr applied to a tiled C (tile list incl. empty tiles from the oracle,
        # whose csr2tile halves are themselves pinned against the reference above)
        oA, oB = orc.csr2tile_row_major(m, n, rp, ci, v), orc.csr2tile_col_major(m, n, rp, ci, v)
        tC = orc.ctiles_from_csr(m, n, oA, oB, (rpC, ciC, vC))
        r2, c2, v2 = ref.tile2csr(tC)
        d.update(C2_rowptr=r2, C2_colidx=c2, C2_val=v2)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(f"{name}: m={m} n={n} nnz={len(ci)} numtileA={tA.numtile}" + (f" nnzC={len(d['C_colidx'])}" if m == n else ""))


def make_general_case(name, tm, tn, m, n, rp, ci, v):
    d = dict(m=np.int64(m), n=np.int64(n), rowptr=rp, colidx=ci, val=v, tile_size=np.array([tm, tn], np.int64))
    tA = ref.csr2tile_row_major(m, n, rp, ci, v, tm, tn)
    tB = ref.csr2tile_col_major(m, n, rp, ci, v, tm, tn)
    d.update(tiled_dict("A", tA))
    d.update(tiled_dict("B", tB))
    if m == n:
        A = (rp, ci, v)
        rpC, ciC, vC = ref.spgemm_serialref(A, A, n)
        d.update(C_rowptr=rpC, C_colidx=ciC, C_val=vC)
        oA, oB = orc.csr2tile_row_major(m, n, rp, ci, v, tm, tn), orc.csr2tile_col_major(m, n, rp, ci, v, tn, tm)
        tC = orc.ctiles_from_csr(m, n, oA, oB, (rpC, ciC, vC))
        r2, c2, v2 = ref.tile2csr(tC)
        d.update(C2_rowptr=r2, C2_colidx=c2, C2_val=v2)
    np.savez_compressed(os.path.join(OUT, f"gtile_{name}_{tm}x{tn}.npz"), **d)
    print(f"gtile_{name}_{tm}x{tn}: numtileA={tA.numtile} numtileB={tB.numtile}")


def main():
    assert ref.available(), "build oracle/_ref first (make -C oracle)"
    for nm in ["diagonal", "tridiagonal", "banded", "sparse", "random_0.05", "random_0.1", "random_0.15"]:
        m, n, rp, ci, v = M.read_mtx(FX + nm + "_36x36.mtx")
        make_case("ref_" + nm.replace(".", "p") + "_36x36", m, n, rp, ci, v)
    # the golden masks of bitmask.h (row masks of random_0.1_36x36.csv, MSB = column 0 of 36)
    txt = open(FX + "bitmask.h").read()
    body = txt[txt.index("{") + 1: txt.index("}")]
    masks = np.array([int(x, 16) for x in re.findall(r"0x[0-9A-Fa-f]+", body)], dtype=np.uint64)
    assert masks.size == 36
    np.savez_compressed(os.path.join(OUT, "ref_bitmask_h_random_0p1.npz"), bitmask=masks)
    # small synthetic cases through the same unmodified reference functions
    make_case("ref_lap2d_32", *M.lap2d(32))
    make_case("ref_rect_100x70", *M.random_sparse(100, 70, 0.05, seed=3))
    make_case("ref_rect_37x129", *M.random_sparse(37, 129, 0.1, seed=4))
    make_case("ref_blockfem_50", *M.blockfem(50))
    make_case("ref_rmat_s8", *M.rmat(8, 8))
    make_case("ref_stencil27_6", *M.stencil27(6))
    # the fork's runtime tile sizes
    for tm, tn in [(32, 32), (16, 32), (32, 16), (48, 64), (64, 64), (128, 128)]:
        make_general_case("lap2d_24", tm, tn, *M.lap2d(24))
        make_general_case("rand_150", tm, tn, *M.random_sparse(150, 150, 0.04, seed=5))
    make_general_case("rect_90x210", 32, 32, *M.random_sparse(90, 210, 0.05, seed=6))
    make_general_case("rect_90x210", 48, 16, *M.random_sparse(90, 210, 0.05, seed=6))
    make_general_case("blockfem_30", 32, 32, *M.blockfem(30))


if __name__ == "__main__":
    main()
