"""General tile sizes on the GPU: csr2tile, SpGEMM steps 1-3 and tile2csr for tiles of
tile_size_m x tile_size_n (A) / tile_size_n x tile_size_m (B) / tile_size_m x tile_size_m (C), through the C ABI
(include/tilespgemm.h Part 3 and the drop-in entry points), against the oracle -- which is pinned to the reference's own
csr2tile / tile2csr at these sizes (tests/test_oracle_vs_ref.py, tests/golden/gtile_*.npz). Structure arrays bit-exact; values
bit-exact for the driver's k % 10 values and within 1e-12 otherwise (the kernels sum in the serial SPA's order with fma())."""
import numpy as np
import pytest
import torch

from conftest import assert_tiled_equal, general_tile_golden_cases, load_golden
from oracle import oracle as orc
from spgemm_b200 import api, matrices as M

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]

VAL_RTOL = 1e-12
B_FIELDS = ("tile_ptr", "tile_columnidx", "tile_nnz", "val", "col", "ptr", "mask", "csc_tile_ptr", "csc_tile_rowidx")
TILE_SIZES = [(16, 16), (32, 32), (16, 32), (32, 16), (48, 64), (64, 64), (128, 128), (128, 16)]

CASES = {
    "lap2d_48": lambda: M.lap2d(48),
    "lap2d_33x17": lambda: M.lap2d(33, 17),
    "stencil27_9": lambda: M.stencil27(9),
    "stencil27_20x7x5": lambda: M.stencil27(20, 7, 5),
    "blockfem_120": lambda: M.blockfem(120),
    "rmat_s10": lambda: M.rmat(10, 8, seed=5),
    "rand_ragged_203": lambda: M.random_sparse(203, 203, 0.03, seed=11),
    "full_48": lambda: M.random_sparse(48, 48, 5.0, seed=15),
    "single_entry": lambda: (20, 20, np.array([0] * 6 + [1] * 15, np.int32), np.array([17], np.int32), np.array([3.0])),
    "empty": lambda: (33, 33, np.zeros(34, np.int32), np.zeros(0, np.int32), np.zeros(0)),
}


@pytest.fixture(scope="module", autouse=True)
def _init():
    api.init(0)
    yield


def run_device_case(tm, tn, A, B=None, exact=True, what=""):
    """Device-resident API: csr2tile x2, steps 1-3, tile2csr -- every array against the oracle. Returns the stats."""
    m, k, rpA, ciA, vA = A
    k2, n, rpB, ciB, vB = B if B is not None else A
    dA = api.DeviceCSR.upload(m, k, rpA, ciA, vA)
    dB = dA if B is None else api.DeviceCSR.upload(k2, n, rpB, ciB, vB)
    tA = api.gtile_csr2tile(dA, False, tm, tn)           # tiles of A: tile_size_m x tile_size_n
    tB = api.gtile_csr2tile(dB, True, tn, tm)            # tiles of B: tile_size_n x tile_size_m (src/main.cu:84)
    oA, oB = orc.csr2tile_row_major(m, k, rpA, ciA, vA, tm, tn), orc.csr2tile_col_major(k2, n, rpB, ciB, vB, tn, tm)
    assert_tiled_equal(tA.download(), oA, f"{what} A {tm}x{tn}")
    assert_tiled_equal(tB.download(), oB, f"{what} B {tn}x{tm}", fields=B_FIELDS)
    tC, st = api.gtile_spgemm(tA, tB)
    csrC = orc.spgemm_spa((rpA, ciA, vA), (rpB, ciB, vB), n)
    oC = orc.ctiles_from_csr(m, n, oA, oB, csrC)
    assert (tC.tile_rows, tC.tile_cols) == (tm, tm)
    assert_tiled_equal(tC.download(), oC, f"{what} C {tm}x{tm}", val_rtol=0.0 if exact else VAL_RTOL)
    assert st["numblkC"] == oC.numtile and st["nnzC"] == oC.nnz and st["pairs"] == int(orc.tilerow_weights(oA, oB).sum())
    assert st["launches"] > 0
    dC = api.gtile_tile2csr(tC)
    r, c, vv = dC.download()
    assert np.array_equal(r, csrC[0]) and np.array_equal(c, csrC[1])
    if exact:
        assert np.array_equal(vv, csrC[2])
    else:
        assert np.allclose(vv, csrC[2], rtol=VAL_RTOL, atol=0.0)
    for o in (dC, tC, tA, tB, dA) + ((dB,) if B is not None else ()):
        o.free()
    return st


@pytest.mark.parametrize("tile", TILE_SIZES, ids=lambda t: f"{t[0]}x{t[1]}")
@pytest.mark.parametrize("name", sorted(CASES))
def test_general_tiles_device_api(name, tile):
    m, n, rp, ci, _ = CASES[name]()
    v = M.set_values(len(ci), "mod10")   # the driver's value[k] = k % 10 (src/main.cu:111-112): every sum is an exact integer
    run_device_case(tile[0], tile[1], (m, n, rp, ci, v), what=name)


@pytest.mark.parametrize("tile", [(32, 32), (48, 16), (16, 64)], ids=lambda t: f"{t[0]}x{t[1]}")
def test_general_tiles_hashed_values(tile):
    m, n, rp, ci, _ = M.stencil27(11, 9, 7)
    run_device_case(tile[0], tile[1], (m, n, rp, ci, M.set_values(len(ci), "hash")), exact=False, what="hash")


@pytest.mark.parametrize("tile", [(32, 32), (64, 16), (16, 48)], ids=lambda t: f"{t[0]}x{t[1]}")
def test_general_tiles_rectangular_product(tile):
    A = M.random_sparse(170, 300, 0.05, seed=21)
    B = M.random_sparse(300, 145, 0.06, seed=22)
    run_device_case(tile[0], tile[1], A, B, what="rect")


@pytest.mark.parametrize("mode", ["gather", "dense"])
@pytest.mark.parametrize("tn", [16, 32, 48, 64])
@pytest.mark.parametrize("name", ["blockfem_120", "stencil27_20x7x5", "full_48", "rmat_s10", "rand_ragged_203", "single_entry"])
def test_general_tiles_numeric_kernels_32_row_tiles(monkeypatch, name, tn, mode):
    """C tiles of 32 x 32 have two numeric kernels: the gather (thread per C nonzero) and the dense accumulator in
    registers (warp per C tile, lane = row; B's tile expanded in shared memory). Same C, bit for bit, from either."""
    monkeypatch.setenv("TSG_GT_NUMERIC", mode)
    m, n, rp, ci, _ = CASES[name]()
    for values, exact in (("mod10", True), ("hash", False)):
        st = run_device_case(32, tn, (m, n, rp, ci, M.set_values(len(ci), values)), exact=exact, what=f"{name}/{mode}")
        assert (st["tiles_dense"] > 0) == (mode == "dense" and st["nnzC"] > 0), st


def test_general_tiles_dense32_selected_for_well_filled_tiles(monkeypatch):
    monkeypatch.delenv("TSG_GT_NUMERIC", raising=False)
    m, n, rp, ci, _ = M.blockfem(400)
    v = M.set_values(len(ci), "mod10")
    assert run_device_case(32, 32, (m, n, rp, ci, v), what="blockfem auto")["tiles_dense"] > 0
    m, n, rp, ci, _ = M.rmat(11, 4, seed=7)
    v = M.set_values(len(ci), "mod10")
    assert run_device_case(32, 32, (m, n, rp, ci, v), what="rmat auto")["tiles_dense"] == 0


@pytest.mark.parametrize("name", general_tile_golden_cases())
def test_general_tiles_vs_reference_golden(name):
    """The arrays the REFERENCE's csr2tile_row_major / csr2tile_col_major(matrix, tile_size_m, tile_size_n) produce
    (tests/golden/gtile_*.npz, generated by make_golden.py from oracle/_ref)."""
    g = load_golden(name)
    m, n, rp, ci, v = int(g["m"]), int(g["n"]), g["rowptr"], g["colidx"], g["val"]
    tm, tn = (int(x) for x in g["tile_size"])
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    for col_major, prefix, (tr, tc) in ((False, "A", (tm, tn)), (True, "B", (tn, tm))):
        t = api.gtile_csr2tile(d, col_major, tr, tc)
        exp = {k[2:]: val for k, val in g.items() if k.startswith(prefix + "_")}
        dm, dn, tilem, tilen, numtile, nnz = (int(x) for x in exp.pop("dims"))
        exp.update(m=dm, n=dn, tilem=tilem, tilen=tilen, numtile=numtile, nnz=nnz)
        assert_tiled_equal(t.download(), exp, f"{name} {prefix}", fields=B_FIELDS if col_major else B_FIELDS[:7] + ("tile_rowidx",))
        t.free()
    d.free()


@pytest.mark.parametrize("name", ["lap2d_48", "stencil27_20x7x5", "blockfem_120", "rmat_s10", "full_48", "empty"])
def test_general_path_at_16x16_equals_the_tuned_path(name):
    """At 16 x 16 the general-tile kernels and the tuned ones must produce the same arrays, bit for bit."""
    m, n, rp, ci, _ = CASES[name]()
    v = M.set_values(len(ci), "hash")
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    tA, tB = api.csr2tile(d, False), api.csr2tile(d, True)
    gA, gB = api.gtile_csr2tile(d, False, 16, 16), api.gtile_csr2tile(d, True, 16, 16)
    assert_tiled_equal(gA.download(), tA.download(), name + " A")
    assert_tiled_equal(gB.download(), tB.download(), name + " B", fields=B_FIELDS)
    tC, st = api.spgemm(tA, tB)
    gC, gst = api.gtile_spgemm(gA, gB)
    assert_tiled_equal(gC.download(), tC.download(), name + " C", val_rtol=VAL_RTOL)   # heavy R-MAT rows: the tuned path's order may differ
    assert (gst["numblkC"], gst["nnzC"], gst["pairs"]) == (st["numblkC"], st["nnzC"], st["pairs"])
    for o in (tC, gC, tA, tB, gA, gB, d):
        o.free()


@pytest.mark.parametrize("tile", [(32, 32), (48, 16)], ids=lambda t: f"{t[0]}x{t[1]}")
def test_drop_in_entry_points_general_tiles(tile):
    """The reference-named calls with the fork's runtime tile sizes, used the way src/main.cu uses them:
    csr2tile_row_major(A, m, n), csr2tile_col_major(B, m, n), tilespgemm(..., m, n), tile2csr(C, m, m) (:168,191,261,327)."""
    tm, tn = tile
    m, n, rp, ci, _ = M.stencil27(9, 8, 7)
    v = M.set_values(len(ci), "mod10")
    A = api.HostMatrix.from_csr(m, n, rp, ci, v)
    B = api.HostMatrix().alias_csr_of(A)
    api.csr2tile_row_major(A, tm, tn)
    api.csr2tile_col_major(B, tm, tn)
    oA, oB = orc.csr2tile_row_major(m, n, rp, ci, v, tm, tn), orc.csr2tile_col_major(m, n, rp, ci, v, tn, tm)
    assert_tiled_equal(A.tiles(tm, tn), oA, "drop-in A")
    assert_tiled_equal(B.tiles(tn, tm), oB, "drop-in B", fields=B_FIELDS)
    nnzCub = orc.nnzcub(ci, rp)
    Cm, info = api.tilespgemm(A, B, nnzCub, tm, tn)
    csrC = orc.spgemm_spa((rp, ci, v), (rp, ci, v), n)
    oC = orc.ctiles_from_csr(m, n, oA, oB, csrC)
    assert_tiled_equal(Cm.tiles(tm, tm), oC, "drop-in C")
    assert info["nnzC_computed"] == oC.nnz and info["compression_rate"] == pytest.approx(nnzCub / oC.nnz)
    assert info["time_tile"] > 0 and info["gflops_tile"] == pytest.approx(2.0 * nnzCub / (info["time_tile"] * 1e6))
    api.tile2csr(Cm, tm, tm)
    r, c, vv = Cm.csr()
    assert np.array_equal(r, csrC[0]) and np.array_equal(c, csrC[1]) and np.array_equal(vv, csrC[2])
    for x in (A, B, Cm):
        api.matrix_destroy(x)


def test_drop_in_general_tiles_unsorted_csr_and_maskless_upload():
    """Rows in scrambled order with duplicates (what the reference's loader hands over, src/mmio_highlevel.h:593-759) are
    canonicalised on the device and tiled; tiles uploaded without their mask array get it rebuilt."""
    import scipy.sparse as sp
    rng = np.random.default_rng(5)
    m, n, rp, ci, _ = M.random_sparse(120, 120, 0.06, seed=31)
    v = M.set_values(len(ci), "mod10")
    S = sp.csr_matrix((v, ci, rp), shape=(m, n))
    rp2, ci2, v2 = rp.copy(), ci.copy(), v.copy()
    for i in range(m):
        s, e = rp[i], rp[i + 1]
        perm = rng.permutation(e - s)
        ci2[s:e], v2[s:e] = ci[s:e][perm], v[s:e][perm]
    tm, tn = 32, 48
    A = api.HostMatrix.from_csr(m, n, rp2, ci2, v2)
    api.csr2tile_row_major(A, tm, tn)
    assert_tiled_equal(A.tiles(tm, tn), orc.csr2tile_row_major(m, n, S.indptr, S.indices, S.data, tm, tn), "unsorted A")
    B = api.HostMatrix.from_csr(m, n, rp, ci, v)
    api.csr2tile_col_major(B, tm, tn)
    keep = (A.s.mask, B.s.mask)
    null = type(A.s.mask)()
    A.s.mask, B.s.mask = null, null
    try:
        Cm, _ = api.tilespgemm(A, B, orc.nnzcub(ci, rp), tm, tn)
    finally:
        A.s.mask, B.s.mask = keep
    csrC = orc.spgemm_spa((rp, ci, v), (rp, ci, v), n)
    oC = orc.ctiles_from_csr(m, n, orc.csr2tile_row_major(m, n, rp, ci, v, tm, tn), orc.csr2tile_col_major(m, n, rp, ci, v, tn, tm), csrC)
    assert_tiled_equal(Cm.tiles(tm, tm), oC, "C from mask-less A, B")
    for x in (A, B, Cm):
        api.matrix_destroy(x)


def test_general_tiles_errors():
    m, n, rp, ci, v = M.lap2d(8)
    A = api.HostMatrix.from_csr(m, n, rp, ci, v)
    for bad in ((24, 16), (16, 144), (0, 16)):
        with pytest.raises(api.TsgError) as e:
            api.csr2tile_row_major(A, *bad)
        assert e.value.code == 2
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    tA = api.gtile_csr2tile(d, False, 32, 16)
    tB = api.gtile_csr2tile(d, True, 32, 32)        # tiles of B must have as many rows as tiles of A have columns
    with pytest.raises(api.TsgError) as e:
        api.gtile_spgemm(tA, tB)
    assert e.value.code == 2
    with pytest.raises(api.TsgError):
        api.gtile_spgemm(tA, tA)                    # B must be column-major tiled
    with pytest.raises(api.TsgError):
        api.gtile_tile2csr(tB)                      # tile2csr wants row-major storage
    for o in (tA, tB, d):
        o.free()


@pytest.mark.parametrize("case", ["lap2d_1100_16x16", "stencil27_48_32x32", "blockfem_20000_64x32", "rmat_s14_32x32"])
def test_general_tiles_larger(case):
    """Sizes where the radix sorts behind csr2tile and step 1 take two and three 8-bit passes (more than 256 / 65 536 tile
    columns), tiles fill up (block-FEM) and tile-rows are skewed (R-MAT)."""
    gen, tm, tn = {
        "lap2d_1100_16x16": (lambda: M.lap2d(1100), 16, 16),          # 75 625 tile columns: three passes
        "stencil27_48_32x32": (lambda: M.stencil27(48), 32, 32),      # 3 456 tile columns: two passes
        "blockfem_20000_64x32": (lambda: M.blockfem(20000), 64, 32),
        "rmat_s14_32x32": (lambda: M.rmat(14, 8, seed=9), 32, 32),
    }[case]
    m, n, rp, ci, _ = gen()
    v = M.set_values(len(ci), "mod10")
    st = run_device_case(tm, tn, (m, n, rp, ci, v), what=case)
    assert st["ms_total"] > 0 and st["ms_step1"] >= 0 and st["algorithmic_bytes"] > 0
