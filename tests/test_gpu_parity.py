"""Parity of the CUDA path (through the C ABI, spgemm_b200.api -> libtilespgemm_b200.so) with the CPU
oracle and the committed golden vectors. Integer/index/structure arrays are compared bit-exact; FP64
values bit-exact under the driver's value[k] = k % 10 convention (all partial sums are exact integers)
and within 1e-12 max relative error for general positive values (BASELINE.json north_star)."""
import os

import numpy as np
import pytest
import torch

from conftest import assert_tiled_equal, golden_cases, load_golden
from oracle import oracle as orc
from spgemm_b200 import api, matrices as M
from test_golden import golden_tiled

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]

VAL_RTOL = 1e-12  # north_star: FP64 values within a max relative error of 1e-12


@pytest.fixture(scope="module", autouse=True)
def _init():
    # tile-row templates (csrc/rowplans.cu) are worth their launches from ~8 K tile-rows on; the tests want them on small matrices too
    os.environ["TSG_ROWPLANS_MIN_ROWS"] = "64"
    api.init(0)
    yield
    os.environ.pop("TSG_ROWPLANS_MIN_ROWS", None)


def oracle_c(m, n, A, B, nB, tA=None, tB=None):
    rpC, ciC, vC = orc.spgemm_spa(A, B, nB)
    tA = tA or orc.csr2tile_row_major(m, n, *A)
    tB = tB or orc.csr2tile_col_major(len(B[0]) - 1, nB, *B)
    return (rpC, ciC, vC), orc.ctiles_from_csr(m, nB, tA, tB, (rpC, ciC, vC))


SMALL = {
    "lap2d_48": lambda: M.lap2d(48),
    "lap2d_33x17": lambda: M.lap2d(33, 17),
    "stencil27_9": lambda: M.stencil27(9),
    "stencil27_20x7x5": lambda: M.stencil27(20, 7, 5),
    "blockfem_120": lambda: M.blockfem(120),
    "blockfem_band3": lambda: M.blockfem(40, dof=6, band=3),
    "rmat_s10": lambda: M.rmat(10, 8, seed=5),
    "rmat_s12_skewed": lambda: M.rmat(12, 16, seed=6),
    "rand_ragged_203": lambda: M.random_sparse(203, 203, 0.03, seed=11),
    "rand_dense_64": lambda: M.random_sparse(64, 64, 0.7, seed=14),
    "full_48": lambda: M.random_sparse(48, 48, 5.0, seed=15),
    "single_entry": lambda: (20, 20, np.array([0] * 6 + [1] * 15, np.int32), np.array([17], np.int32), np.array([3.0])),
    "empty": lambda: (33, 33, np.zeros(34, np.int32), np.zeros(0, np.int32), np.zeros(0)),
    "one_by_one": lambda: (1, 1, np.array([0, 1], np.int32), np.array([0], np.int32), np.array([2.0])),
}


@pytest.mark.parametrize("name", sorted(SMALL))
def test_csr2tile_matches_oracle(name):
    m, n, rp, ci, v = SMALL[name]()
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    for col_major, fn in ((False, orc.csr2tile_row_major), (True, orc.csr2tile_col_major)):
        t = api.csr2tile(d, col_major)
        assert_tiled_equal(t.download(), fn(m, n, rp, ci, v), f"{name} col_major={col_major}")
        t.free()
    d.free()


@pytest.mark.parametrize("name", golden_cases())
def test_golden_vectors(name):
    """csr2tile (both layouts), transposition, SpGEMM and tile2csr against the reference-generated vectors."""
    g = load_golden(name)
    m, n, rp, ci, v = int(g["m"]), int(g["n"]), g["rowptr"], g["colidx"], g["val"]
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    tA, tB = api.csr2tile(d, False), api.csr2tile(d, True)
    assert_tiled_equal(tA.download(), golden_tiled(g, "A"), name + " A")
    assert_tiled_equal(tB.download(), golden_tiled(g, "B"), name + " B")
    at = api.transpose(d)
    cp, ri, cv = at.download()
    assert np.array_equal(cp, g["T_colptr"]) and np.array_equal(ri, g["T_rowidx"]) and np.array_equal(cv, g["T_val"])
    at.free()
    if "C_rowptr" in g:
        tC, st = api.spgemm(tA, tB)
        csr = api.tile2csr_device(tC)
        r, c, vv = csr.download()
        assert np.array_equal(r, g["C_rowptr"]) and np.array_equal(c, g["C_colidx"])
        err = np.max(np.abs(vv - g["C_val"]) / np.maximum(np.abs(g["C_val"]), 1e-300)) if vv.size else 0.0
        assert err <= VAL_RTOL, err
        assert st["nnzC"] == len(g["C_colidx"])
        csr.free(); tC.free()
    tA.free(); tB.free(); d.free()


@pytest.mark.parametrize("values", ["mod10", "hash"])
@pytest.mark.parametrize("name", sorted(SMALL))
def test_spgemm_matches_oracle(name, values):
    m, n, rp, ci, _ = SMALL[name]()
    v = M.set_values(len(ci), values) if len(ci) else np.zeros(0)
    A = (rp, ci, v)
    csrC, tC_exp = oracle_c(m, n, A, A, n)
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    tA, tB = api.csr2tile(d, False), api.csr2tile(d, True)
    assert api.nnzcub(d, d) == orc.nnzcub(ci, rp)
    tC, st = api.spgemm(tA, tB)
    got = tC.download()
    # structure bit-exact (tile list incl. empty tiles, tile_nnz, Ptr, mask, Col); values exact for mod10
    assert_tiled_equal(got, tC_exp, f"{name}/{values} C", val_rtol=0.0 if values == "mod10" else VAL_RTOL)
    assert st["numblkC"] == tC_exp.numtile and st["nnzC"] == tC_exp.nnz
    assert st["pairs"] == int(orc.tilerow_weights(orc.csr2tile_row_major(m, n, *A), orc.csr2tile_col_major(m, n, *A)).sum())
    csr = api.tile2csr_device(tC)
    r, c, vv = csr.download()
    assert np.array_equal(r, csrC[0]) and np.array_equal(c, csrC[1])
    if values == "mod10":
        assert np.array_equal(vv, csrC[2])
    for o in (csr, tC, tA, tB, d):
        o.free()


def test_aat_mode():
    """-aat 1: B = A^T materialised by matrix_transposition (reference src/main.cu:114-142)."""
    m, n, rp, ci, v = M.rmat(10, 8, seed=9)
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    dT = api.transpose(d)
    cp, ri, cv = dT.download()
    ecp, eri, ecv = orc.transpose(m, n, rp, ci, v)
    assert np.array_equal(cp, ecp) and np.array_equal(ri, eri) and np.array_equal(cv, ecv)
    A, B = (rp, ci, v), (ecp, eri, ecv)
    csrC, tC_exp = oracle_c(m, n, A, B, m)
    tA, tB = api.csr2tile(d, False), api.csr2tile(dT, True)
    tC, _ = api.spgemm(tA, tB)
    assert_tiled_equal(tC.download(), tC_exp, "AAT C")
    for o in (tC, tA, tB, d, dT):
        o.free()


def test_rectangular_general_product():
    """General C = A*B with distinct rectangular A and B (ragged last tiles on every edge)."""
    m, k, rpA, ciA, vA = M.random_sparse(70, 100, 0.05, seed=21)
    _, n, rpB, ciB, vB = M.random_sparse(100, 45, 0.06, seed=22)
    A, B = (rpA, ciA, vA), (rpB, ciB, vB)
    csrC, tC_exp = oracle_c(m, k, A, B, n)
    dA, dB = api.DeviceCSR.upload(m, k, *A), api.DeviceCSR.upload(k, n, *B)
    tA, tB = api.csr2tile(dA, False), api.csr2tile(dB, True)
    tC, _ = api.spgemm(tA, tB)
    assert_tiled_equal(tC.download(), tC_exp, "rect C")
    r, c, vv, _ = api.spgemm_csr_host(m, k, n, A, B)
    assert np.array_equal(r, csrC[0]) and np.array_equal(c, csrC[1]) and np.array_equal(vv, csrC[2])
    for o in (tC, tA, tB, dA, dB):
        o.free()


def test_slabs_concatenate_to_whole():
    """C computed slab by slab (the unit the multi-GPU path distributes) equals C computed whole."""
    m, n, rp, ci, v = M.stencil27(12)
    A = (rp, ci, v)
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    tA, tB = api.csr2tile(d, False), api.csr2tile(d, True)
    oA, oB = orc.csr2tile_row_major(m, n, *A), orc.csr2tile_col_major(m, n, *A)
    w = api.tilerow_weights(tA, tB)
    assert np.array_equal(w, orc.tilerow_weights(oA, oB))
    cuts = [0, 5, 6, 40, tA.tilem]
    for t0, t1 in zip(cuts[:-1], cuts[1:]):
        rows = slice(t0 * 16, min(t1 * 16, m))
        sub = orc.spgemm_spa(A, A, n, rows.start, rows.stop)
        exp = orc.ctiles_from_csr(m, n, oA, oB, sub, t0, t1)
        tC, st = api.spgemm(tA, tB, t0, t1)
        got = tC.download()
        assert got["trow0"] == t0
        assert_tiled_equal(got, exp, f"slab [{t0},{t1})")
        assert st["pairs"] == int(w[t0:t1].sum())
        csr = api.tile2csr_device(tC)
        r, c, vv = csr.download()
        assert np.array_equal(r, sub[0]) and np.array_equal(c, sub[1]) and np.array_equal(vv, sub[2])
        csr.free(); tC.free()
    for o in (tA, tB, d):
        o.free()


def test_drop_in_entry_points():
    """The reference-named calls, used the way src/main.cu uses them (host buffers in and out)."""
    m, n, rp, ci, v = M.lap2d(40)
    A = api.HostMatrix.from_csr(m, n, rp, ci, v)
    B = api.HostMatrix().alias_csr_of(A)        # -aat 0: B aliases A's CSR (main.cu:145-151)
    api.csr2tile_row_major(A, 16, 16)
    api.csr2tile_col_major(B, 16, 16)
    oA, oB = orc.csr2tile_row_major(m, n, rp, ci, v), orc.csr2tile_col_major(m, n, rp, ci, v)
    assert_tiled_equal(A.tiles(), oA, "drop-in A")
    assert_tiled_equal(B.tiles(), oB, "drop-in B")
    nnzCub = orc.nnzcub(ci, rp)
    Cm, info = api.tilespgemm(A, B, nnzCub)
    csrC, tC_exp = oracle_c(m, n, (rp, ci, v), (rp, ci, v), n, oA, oB)
    assert_tiled_equal(Cm.tiles(), tC_exp, "drop-in C")
    assert info["nnzC_computed"] == tC_exp.nnz
    assert info["compression_rate"] == pytest.approx(nnzCub / tC_exp.nnz)
    assert info["time_tile"] > 0 and info["gflops_tile"] == pytest.approx(2.0 * nnzCub / (info["time_tile"] * 1e6))
    api.tile2csr(Cm, 16, 16)
    r, c, vv = Cm.csr()
    assert np.array_equal(r, csrC[0]) and np.array_equal(c, csrC[1]) and np.array_equal(vv, csrC[2])
    cp, ri, cv = api.matrix_transposition(m, n, rp, ci, v)
    e = orc.transpose(m, n, rp, ci, v)
    assert np.array_equal(cp, e[0]) and np.array_equal(ri, e[1]) and np.array_equal(cv, e[2])
    for x in (A, B, Cm):
        api.matrix_destroy(x)


def test_error_paths_fail_loudly():
    m, n, rp, ci, v = M.lap2d(8)
    A = api.HostMatrix.from_csr(m, n, rp, ci, v)
    with pytest.raises(api.TsgError) as e:
        api.csr2tile_row_major(A, 24, 16)           # tile sizes are multiples of 16 (one mask word = 16 columns)
    assert e.value.code == 2
    bad = ci.copy()
    bad[0], bad[1] = bad[1], bad[0]                  # unsorted row
    d = api.DeviceCSR.upload(m, n, rp, bad, v)
    with pytest.raises(api.TsgError) as e:
        api.csr2tile(d, False)
    assert e.value.code == 4
    d.free()
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    tA = api.csr2tile(d, False)
    with pytest.raises(api.TsgError):
        api.spgemm(tA, tA)                           # B must be col-major tiled
    tA.free(); d.free()


def test_config1_lap2d_256_full():
    """BASELINE config 1 at full size against the oracle (the reference's CPU-runnable case)."""
    m, n, rp, ci, v = M.lap2d(256)
    A = (rp, ci, v)
    csrC, tC_exp = oracle_c(m, n, A, A, n)
    r, c, vv, st = api.spgemm_csr_host(m, n, n, A)
    assert st["numblkC"] == 50532 and st["nnzC"] == 846852 and st["pairs"] == 97512
    assert np.array_equal(r, csrC[0]) and np.array_equal(c, csrC[1]) and np.array_equal(vv, csrC[2])
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    tA, tB = api.csr2tile(d, False), api.csr2tile(d, True)
    tC, _ = api.spgemm(tA, tB)
    assert_tiled_equal(tC.download(), tC_exp, "lap2d-256 C")
    for o in (tC, tA, tB, d):
        o.free()


def _expected_rowsums(A, B, nB):
    """C * ones = A * (B * ones), exact in FP64 for the integer-valued k % 10 convention."""
    import scipy.sparse as sp
    mA, mB = len(A[0]) - 1, len(B[0]) - 1
    SA = sp.csr_matrix((A[2], A[1], A[0]), shape=(mA, mB))
    SB = sp.csr_matrix((B[2], B[1], B[0]), shape=(mB, nB))
    return SA @ (SB @ np.ones(nB))


def test_rmat_aat_slabwise_matches_oracle():
    """Config 3 at reduced scale (R-MAT, Graph500 skew, C = A A^T), executed slab by slab with a small pair budget
    so that hub tile-rows get slabs of their own and the heavy (multi-warp) step-1 path runs."""
    m, n, rp, ci, v = M.rmat(13, 16, seed=3)
    A = (rp, ci, v)
    cp, ri, cv = orc.transpose(m, n, rp, ci, v)
    B = (cp, ri, cv)
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    dT = api.transpose(d)
    tA, tB = api.csr2tile(d, False), api.csr2tile(dT, True)
    oA, oB = orc.csr2tile_row_major(m, n, *A), orc.csr2tile_col_major(n, m, *B)
    w = api.tilerow_weights(tA, tB)
    assert w.max() > 2048, "the test must exercise the heavy step-1 path"
    exp_sums = _expected_rowsums(A, B, m)
    whole = orc.spgemm_spa(A, B, m)
    seen = []

    def sink(tC, st):
        t0 = tC.trow0
        t1 = t0 + tC.tilem
        r0, r1 = t0 * 16, min(t1 * 16, m)
        sub = (whole[0][r0:r1 + 1] - whole[0][r0], whole[1][whole[0][r0]:whole[0][r1]], whole[2][whole[0][r0]:whole[0][r1]])
        exp = orc.ctiles_from_csr(m, m, oA, oB, sub, t0, t1)
        assert_tiled_equal(tC.download(), exp, f"rmat slab [{t0},{t1})")
        s, c = api.tile_rowsums(tC)
        assert np.array_equal(s, exp_sums[r0:r1]) and np.array_equal(c, np.diff(whole[0][r0:r1 + 1]))
        seen.append((t0, t1))

    tot, per = api.spgemm_slabs(tA, tB, max_pairs=200000, sink=sink, weights=w)
    assert len(per) > 3 and seen[0][0] == 0 and seen[-1][1] == tA.tilem
    assert tot["nnzC"] == whole[0][-1] and tot["pairs"] == int(w.sum())
    for o in (tA, tB, d, dT):
        o.free()


def test_config2_stencil27_128_full():
    """BASELINE config 2 at full size (2.1M rows, 55.7M nnz, 1.49e9 products): the whole CSR of C against the
    oracle, plus the known sizes and the C*ones = A*(A*ones) property."""
    m, n, rp, ci, v = M.stencil27(128)
    A = (rp, ci, v)
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    tA, tB = api.csr2tile(d, False), api.csr2tile(d, True)
    assert (tA.numtile, tB.numtile) == (3210328, 3210328)
    assert api.nnzcub(d, d) == 1489355288
    tC, st = api.spgemm(tA, tB)
    assert (st["numblkC"], st["nnzC"], st["pairs"]) == (13666504, 254840104, 80858168)
    assert st["algorithmic_bytes"] == 5179140416
    s, c = api.tile_rowsums(tC)
    assert np.array_equal(s, _expected_rowsums(A, A, n))
    csr = api.tile2csr_device(tC)
    r, cc, vv = csr.download()
    er, ec, ev = orc.spgemm_spa(A, A, n)
    assert np.array_equal(c, np.diff(er))
    assert np.array_equal(r, er) and np.array_equal(cc, ec) and np.array_equal(vv, ev)
    for o in (csr, tC, tA, tB, d):
        o.free()


def test_config4_blockfem_full():
    """BASELINE config 4 at full size (2M rows, dense 6x6 blocks): known sizes, size-independent properties, and every
    array of the tiled C (tile list incl. the 40 % empty tiles, tile_nnz, Ptr, mask, Col, Val) against the oracle."""
    m, n, rp, ci, v = M.blockfem(333334)
    A = (rp, ci, v)
    assert (m, len(ci)) == (2000004, 36000000)
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    tA, tB = api.csr2tile(d, False), api.csr2tile(d, True)
    assert tA.numtile == 375001 and api.nnzcub(d, d) == 647999136
    tC, st = api.spgemm(tA, tB)
    assert (st["numblkC"], st["nnzC"], st["pairs"]) == (624999, 59999904, 1124999)
    s, c = api.tile_rowsums(tC)
    assert np.array_equal(s, _expected_rowsums(A, A, n))
    # round trip: tile2csr(C) re-tiled equals C's non-empty tiles
    csr = api.tile2csr_device(tC)
    back = api.csr2tile(csr, False)
    assert back.numtile == 375001 and back.nnz == st["nnzC"]
    assert st["tiles_dense"] > 0 or st["plan_recipes"] > 0, st   # the dense accumulator (or, by default, the recipe plans)
    _, tC_exp = oracle_c(m, n, A, A, n)
    assert_tiled_equal(tC.download(), tC_exp, "blockfem-2M C")
    for o in (back, csr, tC, tA, tB, d):
        o.free()


def test_config3_rmat_s16_aat_properties():
    """Config 3 (R-MAT, Graph500 skew, AA^T) at scale 16, slab-wise: totals and row checksums against
    A*(A^T*ones) and the oracle's per-row counts (the full scale-20 run is a bench workload)."""
    m, n, rp, ci, v = M.rmat(16, 16, seed=1)
    A = (rp, ci, v)
    cp, ri, cv = orc.transpose(m, n, rp, ci, v)
    B = (cp, ri, cv)
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    dT = api.transpose(d)
    tA, tB = api.csr2tile(d, False), api.csr2tile(dT, True)
    exp_sums = _expected_rowsums(A, B, m)
    exp_cnt = np.diff(orc.spgemm_spa(A, B, m)[0])
    sums, cnts = np.zeros(m), np.zeros(m, np.int64)

    def sink(tC, st):
        r0 = tC.trow0 * 16
        s, c = api.tile_rowsums(tC)
        sums[r0:r0 + len(s)] = s
        cnts[r0:r0 + len(c)] = c

    tot, per = api.spgemm_slabs(tA, tB, max_pairs=1 << 26, sink=sink)
    assert tot["nnzC"] == int(exp_cnt.sum()) and len(per) >= 2
    assert np.array_equal(cnts, exp_cnt) and np.array_equal(sums, exp_sums)
    for o in (tA, tB, d, dT):
        o.free()


def test_c_driver_cli(tmp_path):
    """driver/test_b200: the reference's CLI (./test -d 0 -aat X file tile_m tile_n) on a .mtx file and on a
    generated matrix, both -aat modes; its own serial-SPA check must pass and the reference's CSV logs are written."""
    import os
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "driver", "test_b200")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "driver")])  # rebuilt whenever the header or the library changed
    m, n, rp, ci, v = M.random_sparse(120, 120, 0.05, seed=5)
    mtx = tmp_path / "rand120.mtx"
    M.write_mtx(str(mtx), m, n, rp, ci, v)
    env = dict(os.environ, TSG_CSV_DIR=str(tmp_path))
    for args in (["-d", "0", "-aat", "0", str(mtx), "16", "16"], ["-d", "0", "-aat", "1", str(mtx), "16", "16"],
                 ["-d", "0", "-aat", "0", "gen:stencil27:12", "16", "16"]):
        out = subprocess.run([exe] + args, capture_output=True, text=True, env=env, timeout=120)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "[PASSED]" in out.stdout and "CUDA  TileSpGEMM runtime is" in out.stdout
    rows = open(tmp_path / "results_tile.csv").read().strip().splitlines()
    assert len(rows) >= 3 and rows[0].split(",")[1:4] == ["120", "120", str(len(ci))]
    for name in ("step_runtime.csv", "mem-cost.csv", "preprocessing.csv"):
        assert len(open(tmp_path / name).read().strip().splitlines()) == 3
    for name in ("results_tile.csv", "step_runtime.csv", "mem-cost.csv", "preprocessing.csv"):
        os.remove(tmp_path / name)
    # the fork's runtime tile sizes (tiles of A 32 x 48, of B 48 x 32, of C 32 x 32): the general-tile path, same check
    for tm, tn in (("32", "32"), ("32", "48")):
        out = subprocess.run([exe, "-d", "0", "-aat", "1", str(mtx), tm, tn], capture_output=True, text=True, env=env, timeout=120)
        assert out.returncode == 0 and "[PASSED]" in out.stdout and f"the tile_size_n = {tn}" in out.stdout, out.stdout + out.stderr
    bad = subprocess.run([exe, "-d", "0", "-aat", "0", str(mtx), "24", "16"], capture_output=True, text=True, env=env, timeout=60)
    assert bad.returncode != 0  # unsupported tile size: the driver exits non-zero instead of printing garbage
    # the reference's loader keeps file order and duplicates (TSG_MTX_RAW=1 does the same): the drop-in csr2tile_* canonicalise
    scr = tmp_path / "scrambled.mtx"
    rng = np.random.default_rng(2)
    S = [(r + 1, c + 1) for r in range(m) for c in ci[rp[r]:rp[r + 1]]]
    rng.shuffle(S)
    S += S[:25]
    with open(scr, "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate pattern general\n{m} {n} {len(S)}\n" + "".join(f"{r} {c}\n" for r, c in S))
    out = subprocess.run([exe, "-d", "0", "-aat", "0", str(scr), "16", "16"], capture_output=True, text=True,
                         env=dict(env, TSG_MTX_RAW="1", TSG_MTX_CACHE="0"), timeout=120)
    assert out.returncode == 0 and "[PASSED]" in out.stdout and "canonicalised on the device" in out.stderr, out.stdout + out.stderr
    # -aat 2 A B: the general product of two files (the CLI shape of the reference's cuSPARSE harness)
    mb, nb, rpb, cib, vb = M.random_sparse(120, 75, 0.07, seed=8)
    mtxb = tmp_path / "rand120x75.mtx"
    M.write_mtx(str(mtxb), mb, nb, rpb, cib, vb)
    out = subprocess.run([exe, "-d", "0", "-aat", "2", str(mtx), str(mtxb), "16", "16"], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and "[PASSED]" in out.stdout and "input matrix B: ( 120, 75 )" in out.stdout, out.stdout + out.stderr
    # -slabs: the product slab by slab through tsg_spgemm_slabs (what configs 3 and 5 need), both modes
    for args in (["-d", "0", "-aat", "1", "gen:rmat:12:16", "16", "16", "-slabs", "100000"],
                 ["-d", "0", "-aat", "0", "gen:stencil27:12", "16", "16", "-slabs", "2000"]):
        out = subprocess.run([exe] + args, capture_output=True, text=True, env=env, timeout=180)
        assert out.returncode == 0 and "[PASSED]" in out.stdout, out.stdout + out.stderr
        nslabs = int([ln for ln in out.stdout.splitlines() if ln.startswith("slabs = ")][0].split()[2])
        assert nslabs >= 3, out.stdout


def test_hypersparse_rmat_a2_matches_oracle():
    """Config 5 in miniature: R-MAT with mild skew, C = A^2. Tiles hold ~1 nonzero, so most listed C tiles are EMPTY
    (tile-level hit, element-level miss) and must still be listed with Ptr = mask = 0; exercises the thread-per-tile
    symbolic (k_step2_thread) and the 1024-thread step-1 path."""
    m, n, rp, ci, v = M.rmat(16, 2, a=0.30, b=0.25, c=0.25, d=0.20, seed=2)
    A = (rp, ci, v)
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    tA, tB = api.csr2tile(d, False), api.csr2tile(d, True)
    assert tA.nnz <= 2 * tA.numtile                                  # hypersparse tiles
    csrC, tC_exp = oracle_c(m, n, A, A, n)
    empty = int((np.diff(tC_exp.tile_nnz) == 0).sum())
    assert empty > tC_exp.numtile // 2, (empty, tC_exp.numtile)       # most C tiles are empty
    tC, st = api.spgemm(tA, tB)
    assert st["pairs"] <= 2 * st["numblkC"]
    assert_tiled_equal(tC.download(), tC_exp, "hypersparse C")
    csr = api.tile2csr_device(tC)
    r, c, vv = csr.download()
    assert np.array_equal(r, csrC[0]) and np.array_equal(c, csrC[1]) and np.array_equal(vv, csrC[2])
    for o in (csr, tC, tA, tB, d):
        o.free()


def _pinned_out(rows, cap):
    return (torch.empty(rows + 1, dtype=torch.int32).pin_memory(), torch.empty(max(cap, 1), dtype=torch.int32).pin_memory(),
            torch.empty(max(cap, 1), dtype=torch.float64).pin_memory())


@pytest.mark.parametrize("nslabs", [0, 1, 3, 7, 64])
def test_overlapped_to_host_equals_oracle_csr(nslabs):
    """tsg_spgemm_to_host: C leaves the device slab by slab on a second stream; the CSR that lands in the
    caller's pinned buffers must be the serial SPA's, whatever the number of slabs (64 > tile-rows: one per row)."""
    m, n, rp, ci, _ = M.stencil27(11, 9, 7)
    v = M.set_values(len(ci), "mod10")
    A = (rp, ci, v)
    exp = orc.spgemm_spa(A, A, n)
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    tA, tB = api.csr2tile(d, False), api.csr2tile(d, True)
    out = _pinned_out(m, len(exp[1]))
    for _ in range(2):  # second pass reuses the landing buffers
        for t in out:
            t.fill_(-1)
        nnz, st = api.spgemm_to_host(tA, tB, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), len(exp[1]), nslabs=nslabs)
        assert nnz == len(exp[1]) == st["nnzC"]
        assert np.array_equal(out[0].numpy(), exp[0])
        assert np.array_equal(out[1].numpy()[:nnz], exp[1]) and np.array_equal(out[2].numpy()[:nnz], exp[2])
    # a sub-range of tile-rows: row pointers are numbered from the range's first row
    t0, t1 = 3, tA.tilem - 2
    sub = orc.spgemm_spa(A, A, n, t0 * 16, min(t1 * 16, m))
    nnz, _ = api.spgemm_to_host(tA, tB, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), len(exp[1]), nslabs=nslabs,
                                trow0=t0, trow1=t1)
    assert nnz == len(sub[1]) and np.array_equal(out[0].numpy()[:len(sub[0])], sub[0])
    assert np.array_equal(out[1].numpy()[:nnz], sub[1]) and np.array_equal(out[2].numpy()[:nnz], sub[2])
    for o in (tA, tB, d):
        o.free()


def test_overlapped_host_into_small_cases_and_capacity_error():
    """tsg_spgemm_csr_host_into on every small case (empty, ragged, 1x1, AA^T), pageable numpy outputs included;
    too small a capacity fails loudly and names the size needed."""
    for name in sorted(SMALL):
        m, n, rp, ci, _ = SMALL[name]()
        v = M.set_values(len(ci), "mod10") if len(ci) else np.zeros(0)
        exp = orc.spgemm_spa((rp, ci, v), (rp, ci, v), n)
        out = (np.full(m + 1, -1, np.int32), np.full(max(len(exp[1]), 1), -1, np.int32), np.full(max(len(exp[1]), 1), -1.0))
        nnz, _ = api.spgemm_csr_host_into(m, n, n, (rp, ci, v), out)
        assert nnz == len(exp[1]), name
        assert np.array_equal(out[0], exp[0]) and np.array_equal(out[1][:nnz], exp[1]) and np.array_equal(out[2][:nnz], exp[2]), name
    m, n, rp, ci, v = M.rmat(10, 8, seed=9)
    ecp, eri, ecv = orc.transpose(m, n, rp, ci, v)
    exp = orc.spgemm_spa((rp, ci, v), (ecp, eri, ecv), m)
    out = (np.zeros(m + 1, np.int32), np.zeros(len(exp[1]), np.int32), np.zeros(len(exp[1])))
    nnz, _ = api.spgemm_csr_host_into(m, n, m, (rp, ci, v), out, aat=True)
    assert nnz == len(exp[1]) and np.array_equal(out[0], exp[0]) and np.array_equal(out[1], exp[1])
    assert np.allclose(out[2], exp[2], rtol=VAL_RTOL, atol=0)
    small = (np.zeros(m + 1, np.int32), np.zeros(len(exp[1]) - 1, np.int32), np.zeros(len(exp[1]) - 1))
    with pytest.raises(api.TsgError) as e:
        api.spgemm_csr_host_into(m, n, m, (rp, ci, v), small, aat=True)
    assert e.value.code == 5 and str(len(exp[1])) in str(e.value)  # TSG_ERR_NOMEM, "C has <nnz> entries"
    nnz, _ = api.spgemm_csr_host_into(m, n, m, (rp, ci, v), out, aat=True)  # and the library is usable afterwards
    assert nnz == len(exp[1]) and np.array_equal(out[1], exp[1])


def test_drop_in_tiles_without_mask_arrays():
    """An SMatrix whose `mask` member is NULL (a caller that tiled A and B itself and kept only Ptr/Col/Val):
    the library rebuilds the row masks on the device, steps 1-3 read them."""
    import ctypes as C
    m, n, rp, ci, _ = M.stencil27(7, 6, 5)
    v = M.set_values(len(ci), "mod10")
    A = api.HostMatrix.from_csr(m, n, rp, ci, v)
    B = api.HostMatrix().alias_csr_of(A)
    api.csr2tile_row_major(A, 16, 16)
    api.csr2tile_col_major(B, 16, 16)
    keep = (A.s.mask, B.s.mask)
    null = type(A.s.mask)()
    A.s.mask, B.s.mask = null, null
    try:
        Cm, _ = api.tilespgemm(A, B, orc.nnzcub(ci, rp))
    finally:
        A.s.mask, B.s.mask = keep       # matrix_destroy frees them
    _, tC_exp = oracle_c(m, n, (rp, ci, v), (rp, ci, v), n)
    assert_tiled_equal(Cm.tiles(), tC_exp, "C from mask-less A, B")
    for x in (A, B, Cm):
        api.matrix_destroy(x)


# ---------------------------------------------------------------------------------------------------------------
# Step 3 picks the accumulator per C tile (dense registers / sparse shared memory / lane-per-nonzero gather):
# whatever it picks, C is the serial SPA's.
# ---------------------------------------------------------------------------------------------------------------
def _mixed_matrix():
    """Block-diagonal [block-FEM | 27-point stencil | R-MAT]: well-filled tiles, sparse tiles and hypersparse tiles
    with a hub tile-row in ONE matrix, so that one call exercises all three numeric kernels."""
    import scipy.sparse as sp
    parts = []
    for gen in (lambda: M.blockfem(60), lambda: M.stencil27(9, 8, 7), lambda: M.rmat(10, 6, seed=4)):
        m, n, rp, ci, v = gen()
        parts.append(sp.csr_matrix((np.ones(len(ci)), ci, rp), shape=(m, n)))
    S = sp.block_diag(parts, format="csr")
    S.sort_indices()
    return S.shape[0], S.shape[1], S.indptr.astype(np.int32), S.indices.astype(np.int32)


@pytest.mark.parametrize("mode", ["auto", "rows", "gather", "dense", "dmma"])
@pytest.mark.parametrize("values", ["mod10", "hash"])
def test_numeric_kernel_selection(monkeypatch, mode, values):
    m, n, rp, ci = _mixed_matrix()
    v = M.set_values(len(ci), values)
    A = (rp, ci, v)
    _, tC_exp = oracle_c(m, n, A, A, n)
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    tA, tB = api.csr2tile(d, False), api.csr2tile(d, True)
    monkeypatch.setenv("TSG_STEP3", mode)
    monkeypatch.setenv("TSG_PLANS", "0")              # the generic numeric kernels are what this test is about
    if mode == "auto":
        monkeypatch.setenv("TSG_ROWS_MIN_FILL", "4")  # let the small R-MAT block split between rows and gather
    tC, st = api.spgemm(tA, tB)
    exact = values == "mod10"
    assert_tiled_equal(tC.download(), tC_exp, f"mixed/{mode}/{values}", val_rtol=0.0 if exact else VAL_RTOL)
    if mode == "auto":
        assert st["tiles_dense"] > 0 and st["rows_staged"] > 0, st
    elif mode == "rows":  # forced, but a tile-row still has to fit the shared-memory budget (the R-MAT hub row does not)
        assert st["rows_staged"] > st["rows_gather"] and st["tiles_dense"] == 0, st
    elif mode == "gather":
        assert st["rows_gather"] > 0 and st["tiles_dense"] == 0 and st["rows_staged"] == 0, st
    else:
        assert st["tiles_dense"] == int((np.diff(tC_exp.tile_nnz) > 0).sum()) and st["rows_staged"] == st["rows_gather"] == 0, st
    for o in (tC, tA, tB, d):
        o.free()


def test_numeric_rows_too_wide_for_shared_memory_fall_back(monkeypatch):
    """A shared-memory budget smaller than most tile-rows: those rows take the gather, the rest stay staged."""
    m, n, rp, ci, _ = M.stencil27(12, 9, 8)
    v = M.set_values(len(ci), "mod10")
    A = (rp, ci, v)
    _, tC_exp = oracle_c(m, n, A, A, n)
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    tA, tB = api.csr2tile(d, False), api.csr2tile(d, True)
    monkeypatch.setenv("TSG_ROWS_SMEM_KB", "12")
    monkeypatch.setenv("TSG_PLANS", "0")
    tC, st = api.spgemm(tA, tB)
    assert st["rows_gather"] > 0 and st["rows_staged"] > 0, st
    assert_tiled_equal(tC.download(), tC_exp, "smem-capped C")
    for o in (tC, tA, tB, d):
        o.free()


def test_values_reproducible_run_to_run():
    """Same inputs, two runs: bit-identical values (no atomics, fixed summation order) on the light path."""
    m, n, rp, ci, _ = M.stencil27(10)
    v = M.set_values(len(ci), "hash")
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    tA, tB = api.csr2tile(d, False), api.csr2tile(d, True)
    a, _ = api.spgemm(tA, tB)
    b, _ = api.spgemm(tA, tB)
    assert np.array_equal(a.download()["val"], b.download()["val"])
    for o in (a, b, tA, tB, d):
        o.free()


# ---------------------------------------------------------------------------------------------------------------
# Boundary: what enters the library is checked before anything indexes by it; unsorted / duplicate-bearing CSR
# ---------------------------------------------------------------------------------------------------------------
def test_out_of_range_input_fails_before_any_kernel_indexes_by_it():
    """Column indices outside [0, n) and broken row pointers are refused at upload (TSG_ERR_INPUT) -- also on the paths
    that would otherwise write ptr[column] (csr2tile_col_major, matrix_transposition) -- and the library stays usable."""
    m, n, rp, ci, v = M.lap2d(12)
    for bad_col in (n, n + 1000, -1, -2 ** 31):
        bad = ci.copy()
        bad[len(bad) // 2] = bad_col
        with pytest.raises(api.TsgError) as e:
            api.DeviceCSR.upload(m, n, rp, bad, v)
        assert e.value.code == 4
        H = api.HostMatrix.from_csr(m, n, rp, bad, v)
        with pytest.raises(api.TsgError) as e:
            api.csr2tile_col_major(H, 16, 16)
        assert e.value.code == 4
        with pytest.raises(api.TsgError) as e:
            api.matrix_transposition(m, n, rp, bad, v)
        assert e.value.code == 4
    brp = rp.copy()
    brp[5], brp[6] = brp[6], brp[5]                      # not monotone
    with pytest.raises(api.TsgError) as e:
        api.DeviceCSR.upload(m, n, brp, ci, v)
    assert e.value.code == 4
    d = api.DeviceCSR.upload(m, n, rp, ci, v)            # still usable
    t = api.csr2tile(d, True)
    assert_tiled_equal(t.download(), orc.csr2tile_col_major(m, n, rp, ci, v), "after the refused uploads")
    t.free(); d.free()


def _scrambled(seed, dups):
    """A CSR whose rows are in random order, optionally with duplicate entries, and its canonical form (scipy)."""
    import scipy.sparse as sp
    rng = np.random.default_rng(seed)
    m, n, rp, ci, _ = M.random_sparse(150, 170, 0.06, seed=seed)
    v = rng.integers(1, 9, len(ci)).astype(np.float64)
    rows = np.repeat(np.arange(m), np.diff(rp))
    if dups:
        extra = rng.choice(len(ci), len(ci) // 5, replace=False)
        rows, ci, v = np.concatenate([rows, rows[extra]]), np.concatenate([ci, ci[extra]]), np.concatenate([v, v[extra] + 1])
    order = np.lexsort((rng.random(len(ci)), rows))      # rows stay contiguous, entries inside a row shuffled
    rows, ci, v = rows[order], ci[order].astype(np.int32), v[order]
    urp = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=m))]).astype(np.int32)
    S = sp.csr_matrix((v.copy(), ci.copy(), urp.copy()), shape=(m, n))  # scipy sorts in place: keep the scrambled arrays
    S.sum_duplicates()
    S.sort_indices()
    return (m, n, urp, ci, v), (S.indptr.astype(np.int32), S.indices.astype(np.int32), S.data)


@pytest.mark.parametrize("dups", [False, True])
def test_canonicalize_and_drop_in_on_unsorted_csr(dups):
    """The reference's loader neither sorts rows nor merges duplicates (src/mmio_highlevel.h:593-759). The device API
    refuses such input (TSG_ERR_INPUT) and offers tsg_csr_canonicalize; the drop-in csr2tile_* canonicalise by themselves."""
    (m, n, urp, uci, uv), (crp, cci, cv) = _scrambled(31 + dups, dups)
    d = api.DeviceCSR.upload(m, n, urp, uci, uv)
    with pytest.raises(api.TsgError) as e:
        api.csr2tile(d, False)
    assert e.value.code == 4
    c = d.canonicalize("sum")
    rp, ci, v = c.download()
    assert np.array_equal(rp, crp) and np.array_equal(ci, cci) and np.array_equal(v, cv)
    if dups:  # keep-first policy: the first occurrence in the input order
        f = d.canonicalize("first")
        frp, fci, fv = f.download()
        assert np.array_equal(frp, crp) and np.array_equal(fci, cci)
        first = {}
        rows = np.repeat(np.arange(m), np.diff(urp))
        for r_, c_, v_ in zip(rows, uci, uv):
            first.setdefault((r_, c_), v_)
        assert np.array_equal(fv, np.array([first[(r_, c_)] for r_, c_ in zip(np.repeat(np.arange(m), np.diff(crp)), cci)]))
        f.free()
    c.free(); d.free()
    # drop-in entry points on the scrambled matrix: same tiles as the canonical matrix gives
    A = api.HostMatrix.from_csr(m, n, urp, uci, uv)
    B = api.HostMatrix().alias_csr_of(A)
    api.csr2tile_row_major(A, 16, 16)
    api.csr2tile_col_major(B, 16, 16)
    assert_tiled_equal(A.tiles(), orc.csr2tile_row_major(m, n, crp, cci, cv), "drop-in A from unsorted CSR")
    assert_tiled_equal(B.tiles(), orc.csr2tile_col_major(m, n, crp, cci, cv), "drop-in B from unsorted CSR")
    if not dups:
        from oracle import ref
        if ref.available():  # the reference's own csr2tile_col_major sorts inside every tile: identical arrays
            assert_tiled_equal(B.tiles(), ref.csr2tile_col_major(m, n, urp, uci, uv), "drop-in B vs reference on unsorted CSR",
                               fields=("tile_ptr", "tile_columnidx", "tile_nnz", "val", "col", "ptr", "mask", "csc_tile_ptr", "csc_tile_rowidx"))
    for x in (A, B):
        api.matrix_destroy(x)


def test_csr_row_slice_is_a_csr_of_its_own():
    m, n, rp, ci, v = M.stencil27(9, 8, 7)
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    for r0, r1 in ((0, m), (16, 80), (100, 100), (m - 5, m)):
        s_ = d.row_slice(r0, r1)
        srp, sci, sv = s_.download()
        assert np.array_equal(srp, rp[r0:r1 + 1] - rp[r0]) and np.array_equal(sci, ci[rp[r0]:rp[r1]]) and np.array_equal(sv, v[rp[r0]:rp[r1]])
        if r1 > r0:
            t = api.csr2tile(s_, False)
            assert_tiled_equal(t.download(), orc.csr2tile_row_major(r1 - r0, n, srp, sci, sv), f"slice [{r0},{r1})")
            t.free()
        s_.free()
    d.free()


# ---------------------------------------------------------------------------------------------------------------
# Recipe plans (csrc/plans.cu): a fast path of steps 2 and 3 for matrices made of few distinct tiles; same results
# ---------------------------------------------------------------------------------------------------------------
STRUCTURED = {
    "stencil27_12x9x8": lambda: M.stencil27(12, 9, 8),
    "stencil27_17": lambda: M.stencil27(17),          # edge not a multiple of 16: many more patterns and recipes
    "lap2d_70x33": lambda: M.lap2d(70, 33),
    "blockfem_200": lambda: M.blockfem(200),
    "blockfem_band2": lambda: M.blockfem(60, dof=6, band=2),
}


@pytest.mark.parametrize("plans", ["1", "0"])
@pytest.mark.parametrize("values", ["mod10", "hash"])
@pytest.mark.parametrize("name", sorted(STRUCTURED))
def test_recipe_plans_same_result(monkeypatch, name, values, plans):
    monkeypatch.setenv("TSG_PLANS", "2" if plans == "1" else "0")   # 2: also on well-filled tiles (block-FEM), where auto prefers the dense kernel
    m, n, rp, ci, _ = STRUCTURED[name]()
    v = M.set_values(len(ci), values)
    A = (rp, ci, v)
    csrC, tC_exp = oracle_c(m, n, A, A, n)
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    tA, tB = api.csr2tile(d, False), api.csr2tile(d, True)
    tC, st = api.spgemm(tA, tB)
    assert (st["plan_recipes"] > 0) == (plans == "1"), st
    # the plans add every C entry's products in the serial SPA's order: bit-exact for integer values, 1e-12 otherwise
    assert_tiled_equal(tC.download(), tC_exp, f"{name}/{values}/plans={plans}", val_rtol=0.0 if values == "mod10" else VAL_RTOL)
    r, c, vv = (x for x in api.tile2csr_device(tC).download())
    assert np.array_equal(r, csrC[0]) and np.array_equal(c, csrC[1])
    # slabs share the recipe machinery: a sub-range of tile-rows through the plans
    t0, t1 = 1, max(2, tA.tilem - 1)
    sub = orc.spgemm_spa(A, A, n, t0 * 16, min(t1 * 16, m))
    tS, st2 = api.spgemm(tA, tB, t0, t1)
    exp = orc.ctiles_from_csr(m, n, orc.csr2tile_row_major(m, n, *A), orc.csr2tile_col_major(m, n, *A), sub, t0, t1)
    assert_tiled_equal(tS.download(), exp, f"{name} slab, plans={plans}", val_rtol=0.0 if values == "mod10" else VAL_RTOL)
    for o in (tS, tC, tA, tB, d):
        o.free()


def test_recipe_plans_numeric_unstaged_rows(monkeypatch):
    """The plan-driven numeric kernel (CTA per tile-row, A's values staged in shared memory, lane per slot -- a chain of C
    nonzeros) has a branch for tile-rows that do not fit its shared memory (lane per nonzero, nothing staged): both
    give the serial SPA's values, bit for bit the same."""
    monkeypatch.setenv("TSG_PLANS", "2")            # 2 = also on well-filled tiles (this small stencil has > 24 entries per tile)
    monkeypatch.setenv("TSG_PLANS_SMEM_KB", "2")
    m, n, rp, ci, _ = M.stencil27(13, 10, 9)
    v = M.set_values(len(ci), "hash")
    A = (rp, ci, v)
    _, tC_exp = oracle_c(m, n, A, A, n)
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    tA, tB = api.csr2tile(d, False), api.csr2tile(d, True)
    tC, st = api.spgemm(tA, tB)
    assert st["plan_recipes"] > 0, st
    assert_tiled_equal(tC.download(), tC_exp, "plans numeric, unstaged rows", val_rtol=VAL_RTOL)
    monkeypatch.delenv("TSG_PLANS_SMEM_KB", raising=False)
    tD, st2 = api.spgemm(tA, tB)                    # the staged branch adds in the same order: bit-identical values
    assert st2["plan_recipes"] > 0, st2
    assert_tiled_equal(tD.download(), tC_exp, "plans numeric, staged rows", val_rtol=VAL_RTOL)
    assert np.array_equal(tC.download()["val"], tD.download()["val"])
    for o in (tD, tC, tA, tB, d):
        o.free()


@pytest.mark.parametrize("chain", ["1", "16", "4000"])
@pytest.mark.parametrize("density", [0.02, 0.3, 1.0])
def test_recipe_plans_slot_chains(monkeypatch, density, chain):
    """A lane of the plan-driven numeric kernel walks a SLOT: a chain of C nonzeros of one tile packed (longest list
    first, into the emptiest slot) to about TSG_PLANS_CHAIN products. Tiles with one nonzero, full 256-entry tiles (up to
    128 slots), chains of one nonzero each (1) and one chain holding the whole tile (4000) all come out exact."""
    monkeypatch.setenv("TSG_PLANS", "2")
    monkeypatch.setenv("TSG_PLANS_CHAIN", chain)
    import scipy.sparse as sp
    pat = sp.random(16, 16, density=density, random_state=3, format="csr")
    pat.data[:] = 1.0
    S = sp.kron(sp.diags([1.0, 1.0, 1.0], [-1, 0, 2], shape=(9, 9)), pat, format="csr")   # 3 tile diagonals of one pattern
    S.sort_indices()
    m = n = S.shape[0]
    rp, ci = S.indptr.astype(np.int32), S.indices.astype(np.int32)
    v = M.set_values(len(ci), "mod10")
    A = (rp, ci, v)
    _, tC_exp = oracle_c(m, n, A, A, n)
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    tA, tB = api.csr2tile(d, False), api.csr2tile(d, True)
    tC, st = api.spgemm(tA, tB)
    assert st["plan_recipes"] > 0, st
    assert_tiled_equal(tC.download(), tC_exp, f"plans slots, density {density}, chain {chain}", val_rtol=0.0)
    for o in (tC, tA, tB, d):
        o.free()


@pytest.mark.parametrize("mode", ["on", "off", "forced_fail"])
@pytest.mark.parametrize("case", ["stencil27_32x16x16", "lap2d_96", "stencil27_slab"])
def test_tile_row_templates_same_result(monkeypatch, case, mode):
    """Tile-row templates (csrc/rowplans.cu): step 1 runs on one representative per distinct tile-row signature and the
    other tile-rows are instantiated from it (and verified element by element). Same C, bit for bit, with the templates
    on, off, and when the verification fails (forced) and the call is redone without them."""
    if mode == "off":
        monkeypatch.setenv("TSG_ROWPLANS", "0")
    if mode == "forced_fail":
        monkeypatch.setenv("TSG_ROWPLANS_FORCE_FAIL", "1")
    if case == "lap2d_96":
        m, n, rp, ci, _ = M.lap2d(96)
    else:
        m, n, rp, ci, _ = M.stencil27(32, 16, 16)
    v = M.set_values(len(ci), "hash")
    A = (rp, ci, v)
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    tA, tB = api.csr2tile(d, False), api.csr2tile(d, True)
    if case == "stencil27_slab":
        t0, t1 = 37, tA.tilem - 29
        sub = orc.spgemm_spa(A, A, n, t0 * 16, min(t1 * 16, m))
        exp = orc.ctiles_from_csr(m, n, orc.csr2tile_row_major(m, n, *A), orc.csr2tile_col_major(m, n, *A), sub, t0, t1)
        tC, st = api.spgemm(tA, tB, t0, t1)
    else:
        _, exp = oracle_c(m, n, A, A, n)
        tC, st = api.spgemm(tA, tB)
    assert st["plan_recipes"] > 0, st
    if mode == "on":
        assert 0 < st["row_templates"] <= tA.tilem // 4, st
    elif mode == "off":
        assert st["row_templates"] == 0, st
    else:
        assert st["row_templates"] == -1, st
    assert_tiled_equal(tC.download(), exp, f"row templates {case} {mode}", val_rtol=VAL_RTOL)
    for o in (tC, tA, tB, d):
        o.free()


def _banded_tiles(T, offsets, pat_seed, density):
    """T x T tiles, one 16x16 pattern repeated on the given tile diagonals: every interior tile-row is a translate of the next."""
    import scipy.sparse as sp
    pat = sp.random(16, 16, density=density, random_state=pat_seed, format="csr")
    pat.data[:] = 1.0
    S = sp.kron(sp.diags([1.0] * len(offsets), offsets, shape=(T, T)), pat, format="csr")
    S.sort_indices()
    return S.shape[0], S.shape[1], S.indptr.astype(np.int32), S.indices.astype(np.int32)


@pytest.mark.parametrize("case", ["many_A_tiles", "long_B_rows"])
def test_tile_row_templates_wide_rows(case):
    """Templates on tile-rows that are wider than a warp: 70 A tiles per tile-row (more than the 64 whose B ranges the
    instantiation keeps in shared memory, and more than one 32-tile chunk of the step-1 walk) and B tile-rows of 40 tiles
    (second chunk of the per-A-tile walk). General C = A*B with different A and B."""
    T = 400
    if case == "many_A_tiles":
        offA, offB = list(range(-35, 35)), list(range(-10, 10))
    else:
        offA, offB = [-3, -1, 0, 2, 7], list(range(-20, 20))
    m, k, rpA, ciA = _banded_tiles(T, offA, 5, 0.03)
    _, n, rpB, ciB = _banded_tiles(T, offB, 9, 0.04)
    A, B = (rpA, ciA, M.set_values(len(ciA), "mod10")), (rpB, ciB, M.set_values(len(ciB), "hash"))
    _, tC_exp = oracle_c(m, k, A, B, n)
    dA, dB = api.DeviceCSR.upload(m, k, *A), api.DeviceCSR.upload(k, n, *B)
    tA, tB = api.csr2tile(dA, False), api.csr2tile(dB, True)
    tC, st = api.spgemm(tA, tB)
    assert st["plan_recipes"] > 0 and 0 < st["row_templates"] <= T // 4, st
    assert_tiled_equal(tC.download(), tC_exp, f"row templates, {case}", val_rtol=VAL_RTOL)
    for o in (tC, tA, tB, dA, dB):
        o.free()


def test_recipe_plans_fall_back_when_recipes_do_not_repeat(monkeypatch):
    """Few tile patterns but tens of thousands of distinct pair sequences: the recipe table overflows, the fail flag
    comes up and the generic kernels produce the result (stats say -1)."""
    monkeypatch.setenv("TSG_PLANS", "1")
    import scipy.sparse as sp
    rng = np.random.default_rng(7)
    T = 256                                             # tiles per side, n = 4096
    pats = [sp.random(16, 16, density=0.04, random_state=s, format="coo") for s in range(8)]
    tiles = sp.random(T, T, density=0.10, random_state=11, format="coo")   # ~2.6 pairs per C tile, 64 choices each; no heavy tile-row
    rows, cols = [], []
    for I, J in zip(tiles.row, tiles.col):
        pm = pats[rng.integers(0, 8)]
        rows.append(I * 16 + pm.row)
        cols.append(J * 16 + pm.col)
    S = sp.csr_matrix((np.ones(sum(len(r_) for r_ in rows)), (np.concatenate(rows), np.concatenate(cols))), shape=(T * 16, T * 16))
    S.sum_duplicates(); S.sort_indices()
    m = n = T * 16
    rp, ci = S.indptr.astype(np.int32), S.indices.astype(np.int32)
    v = M.set_values(len(ci), "mod10")
    A = (rp, ci, v)
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    tA, tB = api.csr2tile(d, False), api.csr2tile(d, True)
    assert 0 < tA.d.npat <= 16 and 0 < tB.d.npat <= 16
    assert api.tilerow_weights(tA, tB).max() <= 2048, "a heavy tile-row would keep the plans from being attempted at all"
    tC, st = api.spgemm(tA, tB)
    assert st["plan_recipes"] == -1, st
    csr = api.tile2csr_device(tC)
    r, c, vv = csr.download()
    er, ec, ev = orc.spgemm_spa(A, A, n)
    assert np.array_equal(r, er) and np.array_equal(c, ec) and np.array_equal(vv, ev)
    for o in (csr, tC, tA, tB, d):
        o.free()


def test_heavy_tile_rows_with_global_memory_bitmap(monkeypatch):
    """Hub tile-rows whose tile-column window does not fit shared memory (forced here with a 1 KB budget; for real from
    ~29 M columns up) keep their step-1 bitmap in global memory, a few CTAs per SM walking the heavy list: same C."""
    monkeypatch.setenv("TSG_S1_SMEM_KB", "1")
    m, n, rp, ci, v = M.rmat(13, 16, seed=3)
    A = (rp, ci, v)
    cp, ri, cv = orc.transpose(m, n, rp, ci, v)
    B = (cp, ri, cv)
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    dT = api.transpose(d)
    tA, tB = api.csr2tile(d, False), api.csr2tile(dT, True)
    assert api.tilerow_weights(tA, tB).max() > 2048          # heavy tile-rows exist
    csrC, tC_exp = oracle_c(m, n, A, B, m)
    tC, st = api.spgemm(tA, tB)
    assert_tiled_equal(tC.download(), tC_exp, "global-bitmap C")
    for o in (tC, tA, tB, d, dT):
        o.free()
