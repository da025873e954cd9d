"""The restatement (oracle/spa_ref.c) against the reference's own CPU code compiled in place
(oracle/_ref). Runs wherever oracle/_ref exists (this container; the GPU box gets the prebuilt .so)."""
import numpy as np
import pytest

from conftest import assert_tiled_equal
from oracle import oracle as orc, ref
from spgemm_b200 import matrices as M

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (no /root/reference here)")

CASES = {
    "lap2d_48": lambda: M.lap2d(48),
    "lap2d_33x17": lambda: M.lap2d(33, 17),
    "stencil27_7": lambda: M.stencil27(7),
    "blockfem_40": lambda: M.blockfem(40),
    "blockfem_band2": lambda: M.blockfem(30, dof=6, band=2),
    "rmat_s9": lambda: M.rmat(9, 8, seed=7),
    "rand_sq_ragged": lambda: M.random_sparse(203, 203, 0.03, seed=11),
    "rand_rect_wide": lambda: M.random_sparse(50, 333, 0.04, seed=12),
    "rand_rect_tall": lambda: M.random_sparse(333, 50, 0.04, seed=13),
    "single_entry": lambda: (20, 20, np.array([0] * 6 + [1] * 15, np.int32), np.array([17], np.int32), np.array([3.0])),
    "empty": lambda: (33, 33, np.zeros(34, np.int32), np.zeros(0, np.int32), np.zeros(0)),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_csr2tile_and_transpose(name):
    m, n, rp, ci, v = CASES[name]()
    assert_tiled_equal(orc.csr2tile_row_major(m, n, rp, ci, v), ref.csr2tile_row_major(m, n, rp, ci, v), name + " A")
    assert_tiled_equal(orc.csr2tile_col_major(m, n, rp, ci, v), ref.csr2tile_col_major(m, n, rp, ci, v), name + " B")
    got, exp = orc.transpose(m, n, rp, ci, v), ref.matrix_transposition(m, n, rp, ci, v)
    for a, e in zip(got, exp):
        assert np.array_equal(a, e)


@pytest.mark.parametrize("name", [k for k in sorted(CASES) if k not in ("rand_rect_wide", "rand_rect_tall")])
def test_spa_and_tile2csr(name):
    m, n, rp, ci, v = CASES[name]()
    A = (rp, ci, v)
    rpC, ciC, vC = orc.spgemm_spa(A, A, n)
    r1, c1 = ref.spgemm_spa(A, A, n)                 # structure (src/spgemm_serialref_spa_new.h)
    r2, c2, v2 = ref.spgemm_serialref(A, A, n)       # values    (external/cusparse/spgemm_serialref_spa.h)
    assert np.array_equal(rpC, r1) and np.array_equal(ciC, c1)
    assert np.array_equal(rpC, r2) and np.array_equal(ciC, c2) and np.array_equal(vC, v2)
    tA, tB = orc.csr2tile_row_major(m, n, rp, ci, v), orc.csr2tile_col_major(m, n, rp, ci, v)
    tC = orc.ctiles_from_csr(m, n, tA, tB, (rpC, ciC, vC))
    for got in (orc.tile2csr(tC), ref.tile2csr(tC)):
        assert np.array_equal(got[0], rpC) and np.array_equal(got[1], ciC) and np.array_equal(got[2], vC)


@pytest.mark.parametrize("tile", [(32, 32), (16, 48), (64, 32), (128, 128)])
@pytest.mark.parametrize("name", ["lap2d_33x17", "stencil27_7", "blockfem_40", "rmat_s9", "rand_rect_wide", "rand_rect_tall", "single_entry", "empty"])
def test_general_tile_sizes(name, tile):
    """Runtime tile sizes: the restatement against the reference's csr2tile_row_major /
    csr2tile_col_major(matrix, tile_size_m, tile_size_n) and tile2csr(C, tile_size_m, tile_size_m)."""
    tm, tn = tile
    m, n, rp, ci, v = CASES[name]()
    tA, tB = orc.csr2tile_row_major(m, n, rp, ci, v, tm, tn), orc.csr2tile_col_major(m, n, rp, ci, v, tn, tm)
    assert_tiled_equal(tA, ref.csr2tile_row_major(m, n, rp, ci, v, tm, tn), name + " A")
    # tile_rowidx of B: allocated and left zero by the reference (src/csr2tile.h:336-337); not compared
    assert_tiled_equal(tB, ref.csr2tile_col_major(m, n, rp, ci, v, tm, tn), name + " B",
                       fields=("tile_ptr", "tile_columnidx", "tile_nnz", "val", "col", "ptr", "mask", "csc_tile_ptr", "csc_tile_rowidx"))
    if m == n:
        A = (rp, ci, v)
        C = orc.spgemm_spa(A, A, n)
        tC = orc.ctiles_from_csr(m, n, tA, tB, C)
        for got in (orc.tile2csr(tC), ref.tile2csr(tC)):
            assert np.array_equal(got[0], C[0]) and np.array_equal(got[1], C[1]) and np.array_equal(got[2], C[2])


def test_rectangular_product():
    m, k, rpA, ciA, vA = M.random_sparse(70, 100, 0.05, seed=21)
    k2, n, rpB, ciB, vB = M.random_sparse(100, 45, 0.06, seed=22)
    got = orc.spgemm_spa((rpA, ciA, vA), (rpB, ciB, vB), n)
    exp = ref.spgemm_serialref((rpA, ciA, vA), (rpB, ciB, vB), n)
    for a, e in zip(got, exp):
        assert np.array_equal(a, e)
