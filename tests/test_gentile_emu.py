"""General tile sizes, CPU half: spgemm_b200/csrc/gentile.cu compiled as plain C++ with one
emulated thread after another (tests/emu/), driven through the three host-buffer calls the drop-in entry points make for a
tile size other than 16 x 16, against the oracle (itself pinned to the reference's csr2tile / tile2csr at these sizes,
tests/test_oracle_vs_ref.py::test_general_tile_sizes). The GPU half is tests/test_gentile_gpu.py: same cases, real kernels.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_tiled_equal, general_tile_golden_cases, load_golden
from oracle import oracle as orc
from spgemm_b200 import matrices as M
from spgemm_b200.lib import SMatrix, Stats

EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "libgentile_emu.so")
SRC = os.path.join(ROOT, "spgemm_b200", "csrc", "gentile.cu")

TILE_SIZES = [(16, 16), (32, 32), (16, 32), (32, 16), (48, 64), (64, 64), (128, 128), (128, 16)]


@pytest.fixture(scope="module")
def emu():
    deps = [SRC, os.path.join(EMU_DIR, "gentile_emu.cpp"), os.path.join(EMU_DIR, "gentile_emu.h"),
            os.path.join(ROOT, "include", "tilespgemm.h")]
    if not os.path.exists(EMU_SO) or any(os.path.getmtime(EMU_SO) < os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-w", "-mfma", "-I", EMU_DIR, "-x", "c++",
                               os.path.join(EMU_DIR, "gentile_emu.cpp"), "-o", EMU_SO])
    lib = C.CDLL(EMU_SO)
    lib.emu_last_error_string.restype = C.c_char_p
    lib.emu_launches.restype = C.c_longlong
    lib.emu_free.restype = None
    lib.emu_clear_error.restype = None
    return lib


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def _take(p, n, dt):
    if n <= 0 or not p:
        return np.zeros(0, dt)
    return np.ctypeslib.as_array(p, shape=(n,)).astype(dt, copy=True)


class Host:
    """An SMatrix whose CSR arrays are numpy-owned and whose tile arrays are malloc()ed by the code under test."""

    def __init__(self, m=0, n=0, rp=None, ci=None, v=None):
        self.s = SMatrix()
        self.keep = []
        if rp is not None:
            rp, ci, v = np.ascontiguousarray(rp, np.int32), np.ascontiguousarray(ci, np.int32), np.ascontiguousarray(v, np.float64)
            self.keep = [rp, ci, v]
            self.s.m, self.s.n, self.s.nnz = int(m), int(n), int(rp[m])
            self.s.rowpointer, self.s.columnindex, self.s.value = _p(rp, C.c_int), _p(ci, C.c_int), _p(v, C.c_double)

    def tiles(self, tr, tc, col_major=False):
        s = self.s
        nt, nnz = s.numtile, s.nnz
        d = dict(m=s.m, n=s.n, tilem=s.tilem, tilen=s.tilen, numtile=nt, nnz=nnz,
                 tile_ptr=_take(s.tile_ptr, s.tilem + 1, np.int32), tile_columnidx=_take(s.tile_columnidx, nt, np.int32),
                 tile_rowidx=_take(s.tile_rowidx, nt, np.int32), tile_nnz=_take(s.tile_nnz, nt + 1, np.int32),
                 val=_take(s.tile_csr_Value, nnz, np.float64), col=_take(s.tile_csr_Col, nnz, np.uint16),
                 ptr=_take(s.tile_csr_Ptr, nt * tr, np.uint16), mask=_take(s.mask, nt * tr * (tc // 16), np.uint16))
        if col_major:
            d["csc_tile_ptr"] = _take(s.csc_tile_ptr, s.tilen + 1, np.int32)
            d["csc_tile_rowidx"] = _take(s.csc_tile_rowidx, nt, np.int32)
        return d

    def csr(self):
        s = self.s
        return _take(s.rowpointer, s.m + 1, np.int32), _take(s.columnindex, s.nnz, np.int32), _take(s.value, s.nnz, np.float64)

    def free_tiles(self, lib):
        s = self.s
        for f in ("tile_ptr", "tile_columnidx", "tile_rowidx", "tile_nnz", "tile_csr_Value", "tile_csr_Col", "tile_csr_Ptr", "mask",
                  "csc_tile_ptr", "csc_tile_rowidx"):
            p = getattr(s, f)
            if p:
                lib.emu_free(p)


def ok(lib, rc):
    assert rc == 0 and lib.emu_last_error() == 0, (rc, lib.emu_last_error_string().decode())


def run_case(lib, tm, tn, A, B=None, values_exact=True):
    """csr2tile(A), csr2tile(B), steps 1-3, tile2csr for tiles of A tm x tn -- every array against the oracle."""
    m, k, rpA, ciA, vA = A
    k2, n, rpB, ciB, vB = B if B is not None else A
    assert k == k2
    lib.emu_clear_error()
    hA, hB, hC = Host(m, k, rpA, ciA, vA), Host(k2, n, rpB, ciB, vB), Host()
    ok(lib, lib.emu_csr2tile(C.byref(hA.s), tm, tn, 0))
    ok(lib, lib.emu_csr2tile(C.byref(hB.s), tm, tn, 1))
    oA, oB = orc.csr2tile_row_major(m, k, rpA, ciA, vA, tm, tn), orc.csr2tile_col_major(k2, n, rpB, ciB, vB, tn, tm)
    assert_tiled_equal(hA.tiles(tm, tn), oA, f"A {tm}x{tn}")
    assert_tiled_equal(hB.tiles(tn, tm, True), oB, f"B {tn}x{tm}",
                       fields=("tile_ptr", "tile_columnidx", "tile_nnz", "val", "col", "ptr", "mask", "csc_tile_ptr", "csc_tile_rowidx"))
    st = Stats()
    ok(lib, lib.emu_tilespgemm(C.byref(hA.s), C.byref(hB.s), C.byref(hC.s), tm, tn, C.byref(st)))
    csrC = orc.spgemm_spa((rpA, ciA, vA), (rpB, ciB, vB), n)
    oC = orc.ctiles_from_csr(m, n, oA, oB, csrC)
    assert (oC.tr, oC.tc) == (tm, tm)
    assert_tiled_equal(hC.tiles(tm, tm), oC, f"C {tm}x{tm}", val_rtol=0.0 if values_exact else 1e-12)
    assert st.numblkC == oC.numtile and st.nnzC == oC.nnz and st.pairs == int(orc.tilerow_weights(oA, oB).sum())
    ok(lib, lib.emu_tile2csr(C.byref(hC.s), tm, tm))           # the driver's tile2csr(C, tile_size_m, tile_size_m), src/main.cu:327
    r, c, vv = hC.csr()
    assert np.array_equal(r, csrC[0]) and np.array_equal(c, csrC[1])
    if values_exact:
        assert np.array_equal(vv, csrC[2])
    else:
        assert np.allclose(vv, csrC[2], rtol=1e-12, atol=0)
    for h in (hA, hB, hC):
        h.free_tiles(lib)
    lib.emu_free(hC.s.rowpointer); lib.emu_free(hC.s.columnindex); lib.emu_free(hC.s.value)
    ok(lib, 0)  # the guard zones of every scratch block were intact when it was freed


CASES = {
    "lap2d_20": lambda: M.lap2d(20),
    "lap2d_33x17": lambda: M.lap2d(33, 17),
    "stencil27_6": lambda: M.stencil27(6),
    "stencil27_9x5x4": lambda: M.stencil27(9, 5, 4),
    "blockfem_24": lambda: M.blockfem(24),
    "rmat_s8": lambda: M.rmat(8, 6, seed=3),
    "rand_ragged_203": lambda: M.random_sparse(203, 203, 0.03, seed=11),
    "dense_40": lambda: M.random_sparse(40, 40, 1.0, seed=2),
    "single_entry": lambda: (20, 20, np.array([0] * 6 + [1] * 15, np.int32), np.array([17], np.int32), np.array([3.0])),
    "empty": lambda: (33, 33, np.zeros(34, np.int32), np.zeros(0, np.int32), np.zeros(0)),
}


@pytest.mark.parametrize("tile", TILE_SIZES, ids=lambda t: f"{t[0]}x{t[1]}")
@pytest.mark.parametrize("name", sorted(CASES))
def test_general_tiles_emulated(emu, name, tile):
    m, n, rp, ci, _ = CASES[name]()
    v = M.set_values(len(ci), "mod10")   # the driver's value[k] = k % 10 (src/main.cu:111-112): every sum is an exact integer
    run_case(emu, tile[0], tile[1], (m, n, rp, ci, v))


@pytest.mark.parametrize("tile", [(32, 32), (48, 16), (16, 64)], ids=lambda t: f"{t[0]}x{t[1]}")
def test_general_tiles_emulated_hashed_values(emu, tile):
    """Non-integer values: the sums are taken in the serial SPA's order with fma(), so they are bit-identical too."""
    m, n, rp, ci, _ = M.stencil27(7, 6, 5)
    run_case(emu, tile[0], tile[1], (m, n, rp, ci, M.set_values(len(ci), "hash")))


@pytest.mark.parametrize("tile", [(32, 32), (64, 16), (16, 48)], ids=lambda t: f"{t[0]}x{t[1]}")
def test_general_tiles_emulated_rectangular_product(emu, tile):
    A = M.random_sparse(70, 100, 0.05, seed=21)
    B = M.random_sparse(100, 45, 0.06, seed=22)
    run_case(emu, tile[0], tile[1], A, B)


@pytest.mark.parametrize("name", general_tile_golden_cases())
def test_general_tiles_emulated_vs_reference_golden(emu, name):
    """The arrays the REFERENCE's csr2tile_row_major / csr2tile_col_major produce at these tile sizes (tests/golden/gtile_*)."""
    g = load_golden(name)
    m, n, rp, ci, v = int(g["m"]), int(g["n"]), g["rowptr"], g["colidx"], g["val"]
    tm, tn = (int(x) for x in g["tile_size"])
    emu.emu_clear_error()
    for col_major, prefix, (tr, tc) in ((0, "A", (tm, tn)), (1, "B", (tn, tm))):
        h = Host(m, n, rp, ci, v)
        ok(emu, emu.emu_csr2tile(C.byref(h.s), tm, tn, col_major))
        exp = {k[2:]: val for k, val in g.items() if k.startswith(prefix + "_")}
        dm, dn, tilem, tilen, numtile, nnz = (int(x) for x in exp.pop("dims"))
        exp.update(m=dm, n=dn, tilem=tilem, tilen=tilen, numtile=numtile, nnz=nnz)
        fields = ("tile_ptr", "tile_columnidx", "tile_nnz", "val", "col", "ptr", "mask") + (("csc_tile_ptr", "csc_tile_rowidx") if col_major else ("tile_rowidx",))
        assert_tiled_equal(h.tiles(tr, tc, bool(col_major)), exp, f"{name} {prefix}", fields=fields)
        h.free_tiles(emu)


def test_general_tiles_emulated_errors(emu):
    m, n, rp, ci, v = M.lap2d(8)
    emu.emu_clear_error()
    h = Host(m, n, rp, ci, v)
    assert emu.emu_csr2tile(C.byref(h.s), 24, 16, 0) == 2      # not a multiple of 16
    emu.emu_clear_error()
    assert emu.emu_csr2tile(C.byref(h.s), 16, 144, 0) == 2     # more than 128
    emu.emu_clear_error()
    bad = ci.copy()
    bad[0], bad[1] = bad[1], bad[0]
    h = Host(m, n, rp, bad, v)
    assert emu.emu_csr2tile(C.byref(h.s), 32, 32, 0) == 4      # unsorted row: TSG_ERR_INPUT (the drop-in canonicalises and retries)
    emu.emu_clear_error()


def test_general_tiles_emulated_under_address_sanitizer(tmp_path):
    """The same cases with the emulated build compiled -fsanitize=address: out-of-bounds loads as well as stores, in the
    kernels and in the host orchestration, abort the run (compute-sanitizer is not available on the GPU pool)."""
    import sys
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not asan or not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan.so not installed")
    so = str(tmp_path / "libgentile_emu_asan.so")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-w", "-mfma", "-fsanitize=address", "-fno-omit-frame-pointer",
                           "-I", EMU_DIR, "-x", "c++", os.path.join(EMU_DIR, "gentile_emu.cpp"), "-o", so])
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0")
    out = subprocess.run([sys.executable, os.path.join(EMU_DIR, "run_asan.py"), so], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and "asan cases ok: 83" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
