"""ctypes front-end of the CPU oracle (oracle/spa_ref.c -> oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY. May be imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by spgemm_b200/ (the product).
Each wrapper returns plain numpy arrays; see spa_ref.c for the reference file:line each
function restates.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> None:
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "spa_ref.c")
    stale = force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src)
    have_ref = os.path.isdir("/root/reference/src")
    ref_so = os.path.join(_HERE, "_ref", "libref_cpu.so")
    if stale or (have_ref and not os.path.exists(ref_so)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])


class _Tiled(C.Structure):
    _fields_ = [
        ("m", C.c_int), ("n", C.c_int), ("tilem", C.c_int), ("tilen", C.c_int), ("numtile", C.c_int),
        ("nnz", C.c_int64),
        ("tile_ptr", C.POINTER(C.c_int)), ("tile_columnidx", C.POINTER(C.c_int)),
        ("tile_rowidx", C.POINTER(C.c_int)), ("tile_nnz", C.POINTER(C.c_int64)),
        ("val", C.POINTER(C.c_double)), ("col", C.POINTER(C.c_uint16)),
        ("ptr", C.POINTER(C.c_uint16)), ("mask", C.POINTER(C.c_uint16)),
        ("csc_tile_ptr", C.POINTER(C.c_int)), ("csc_tile_rowidx", C.POINTER(C.c_int)),
        ("tr", C.c_int), ("tc", C.c_int),
    ]


class _Csr(C.Structure):
    _fields_ = [("m", C.c_int), ("n", C.c_int), ("nnz", C.c_int64),
                ("rowptr", C.POINTER(C.c_int64)), ("colidx", C.POINTER(C.c_int)), ("val", C.POINTER(C.c_double))]


def lib():
    global _LIB
    if _LIB is None:
        build()
        _LIB = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        _LIB.orc_nnzcub.restype = C.c_uint64
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


@dataclass
class Tiled:
    """Tiled matrix as numpy arrays (layout: SURVEY.md Appendix A; reference src/common.h:150-172)."""
    m: int
    n: int
    tilem: int
    tilen: int
    numtile: int
    nnz: int
    tile_ptr: np.ndarray
    tile_columnidx: np.ndarray
    tile_rowidx: np.ndarray
    tile_nnz: np.ndarray          # int64 exclusive offsets [numtile+1]
    val: np.ndarray
    col: np.ndarray               # uint16
    ptr: np.ndarray               # uint16 [numtile*tr]
    mask: np.ndarray              # uint16 [numtile*tr*(tc/16)]
    csc_tile_ptr: np.ndarray | None = None
    csc_tile_rowidx: np.ndarray | None = None
    tr: int = 16                  # rows of one tile
    tc: int = 16                  # columns of one tile (general tiles: SURVEY.md 8(f) rank 1)


def _arr(p, n, dtype):
    if n <= 0 or not p:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(p, shape=(n,)).astype(dtype, copy=True)


def _take_tiled(t: _Tiled) -> Tiled:
    nt = t.numtile
    tr, tc = (t.tr or 16), (t.tc or 16)
    out = Tiled(
        m=t.m, n=t.n, tilem=t.tilem, tilen=t.tilen, numtile=nt, nnz=int(t.nnz),
        tile_ptr=_arr(t.tile_ptr, t.tilem + 1, np.int32),
        tile_columnidx=_arr(t.tile_columnidx, nt, np.int32),
        tile_rowidx=_arr(t.tile_rowidx, nt, np.int32),
        tile_nnz=_arr(t.tile_nnz, nt + 1, np.int64),
        val=_arr(t.val, int(t.nnz), np.float64),
        col=_arr(t.col, int(t.nnz), np.uint16),
        ptr=_arr(t.ptr, nt * tr, np.uint16),
        mask=_arr(t.mask, nt * tr * (tc // 16), np.uint16),
        csc_tile_ptr=_arr(t.csc_tile_ptr, t.tilen + 1, np.int32) if t.csc_tile_ptr else None,
        csc_tile_rowidx=_arr(t.csc_tile_rowidx, nt, np.int32) if t.csc_tile_rowidx else None,
        tr=tr, tc=tc,
    )
    lib().orc_tiled_free(C.byref(t))
    return out


def _take_csr(c: _Csr):
    out = (_arr(c.rowptr, c.m + 1, np.int64), _arr(c.colidx, int(c.nnz), np.int32),
           _arr(c.val, int(c.nnz), np.float64))
    lib().orc_csr_free(C.byref(c))
    return out


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def _csr_in(rowptr, colidx, val):
    rp = np.ascontiguousarray(rowptr, dtype=np.int64)
    ci = np.ascontiguousarray(colidx, dtype=np.int32)
    v = np.ascontiguousarray(val, dtype=np.float64)
    return rp, ci, v


def csr2tile_row_major(m, n, rowptr, colidx, val, tr=16, tc=16) -> Tiled:
    """Tiles of tr rows x tc columns (the reference's csr2tile_row_major(A, tile_size_m = tr, tile_size_n = tc))."""
    rp, ci, v = _csr_in(rowptr, colidx, val)
    t = _Tiled()
    rc = lib().orc_csr2tile_row_major_g(int(m), int(n), _p(rp, C.c_int64), _p(ci, C.c_int), _p(v, C.c_double),
                                        int(tr), int(tc), C.byref(t))
    assert rc == 0, f"tile size {tr}x{tc} not representable"
    return _take_tiled(t)


def csr2tile_col_major(m, n, rowptr, colidx, val, tr=16, tc=16) -> Tiled:
    """Tiles of tr rows x tc columns in CSC-tile order. NB the reference's csr2tile_col_major(B, tile_size_m,
    tile_size_n) makes tiles of tile_size_n rows x tile_size_m columns: pass tr = tile_size_n, tc = tile_size_m."""
    rp, ci, v = _csr_in(rowptr, colidx, val)
    t = _Tiled()
    rc = lib().orc_csr2tile_col_major_g(int(m), int(n), _p(rp, C.c_int64), _p(ci, C.c_int), _p(v, C.c_double),
                                        int(tr), int(tc), C.byref(t))
    assert rc == 0, f"tile size {tr}x{tc} not representable"
    return _take_tiled(t)


def transpose(m, n, rowptr, colidx, val):
    """CSR -> CSC, stable (reference src/utils.h:161). Returns (colptr int64, rowidx, cscval)."""
    rp, ci, v = _csr_in(rowptr, colidx, val)
    nnz = int(rp[m])
    colptr = np.zeros(n + 1, dtype=np.int64)
    rowidx = np.zeros(nnz, dtype=np.int32)
    cv = np.zeros(nnz, dtype=np.float64)
    lib().orc_transpose(int(m), int(n), _p(rp, C.c_int64), _p(ci, C.c_int), _p(v, C.c_double),
                        _p(colptr, C.c_int64), _p(rowidx, C.c_int), _p(cv, C.c_double))
    return colptr, rowidx, cv


def nnzcub(colidxA, rowptrB) -> int:
    ci = np.ascontiguousarray(colidxA, dtype=np.int32)
    rp = np.ascontiguousarray(rowptrB, dtype=np.int64)
    return int(lib().orc_nnzcub(C.c_int64(ci.size), _p(ci, C.c_int), _p(rp, C.c_int64)))


def spgemm_spa(A, B, nB, row0=0, row1=None):
    """C = A*B rows [row0,row1) by SPA with values. A, B = (rowptr, colidx, val)."""
    rpA, ciA, vA = _csr_in(*A)
    rpB, ciB, vB = _csr_in(*B)
    mA = rpA.size - 1
    if row1 is None:
        row1 = mA
    c = _Csr()
    rc = lib().orc_spgemm_spa(int(mA), int(nB), _p(rpA, C.c_int64), _p(ciA, C.c_int), _p(vA, C.c_double),
                              _p(rpB, C.c_int64), _p(ciB, C.c_int), _p(vB, C.c_double),
                              int(row0), int(row1), C.byref(c))
    assert rc == 0
    return _take_csr(c)


def ctiles_from_csr(m, n, tA: Tiled, tB: Tiled, csrC, trow0=0, trow1=None) -> Tiled:
    """Tiled C (incl. empty tiles) from CSR(C) and the tile patterns of A and B. C's tiles have A's tile rows x B's
    tile columns (tile_size_m x tile_size_m in the reference's terms)."""
    assert tA.tc == tB.tr, "inner tile dimension: A's tile columns must equal B's tile rows"
    rp, ci, v = _csr_in(*csrC)
    if trow1 is None:
        trow1 = tA.tilem
    pa, ca = np.ascontiguousarray(tA.tile_ptr, np.int32), np.ascontiguousarray(tA.tile_columnidx, np.int32)
    pb, cb = np.ascontiguousarray(tB.tile_ptr, np.int32), np.ascontiguousarray(tB.tile_columnidx, np.int32)
    t = _Tiled()
    rc = lib().orc_ctiles_from_csr_g(int(m), int(n), int(tA.tilem), _p(pa, C.c_int), _p(ca, C.c_int),
                                     int(tB.tilen), _p(pb, C.c_int), _p(cb, C.c_int),
                                     int(trow0), int(trow1),
                                     _p(rp, C.c_int64), _p(ci, C.c_int), _p(v, C.c_double), int(tA.tr), int(tB.tc), C.byref(t))
    assert rc == 0, "C has an entry outside the tile-level product"
    return _take_tiled(t)


def tile2csr(t: Tiled):
    s = _Tiled()
    keep = []

    def put(name, arr, ct, dt):
        a = np.ascontiguousarray(arr, dtype=dt)
        keep.append(a)
        setattr(s, name, _p(a, ct))

    s.m, s.n, s.tilem, s.tilen, s.numtile, s.nnz = t.m, t.n, t.tilem, t.tilen, t.numtile, t.nnz
    s.tr, s.tc = t.tr, t.tc
    put("tile_ptr", t.tile_ptr, C.c_int, np.int32)
    put("tile_columnidx", t.tile_columnidx, C.c_int, np.int32)
    put("tile_nnz", t.tile_nnz, C.c_int64, np.int64)
    put("val", t.val, C.c_double, np.float64)
    put("col", t.col, C.c_uint16, np.uint16)
    put("ptr", t.ptr, C.c_uint16, np.uint16)
    c = _Csr()
    rc = lib().orc_tile2csr(C.byref(s), C.byref(c))
    assert rc == 0
    return _take_csr(c)


def tilerow_weights(tA: Tiled, tB: Tiled) -> np.ndarray:
    pa, ca = np.ascontiguousarray(tA.tile_ptr, np.int32), np.ascontiguousarray(tA.tile_columnidx, np.int32)
    pb = np.ascontiguousarray(tB.tile_ptr, np.int32)
    w = np.zeros(tA.tilem, dtype=np.int64)
    lib().orc_tilerow_weights(int(tA.tilem), _p(pa, C.c_int), _p(ca, C.c_int), _p(pb, C.c_int), _p(w, C.c_int64))
    return w


def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int | None = None) -> int:
    """OpenMP team size of the oracle (None: every host core). torchrun exports OMP_NUM_THREADS=1 to its workers,
    which would time the CPU baseline on one core."""
    lib().orc_set_num_threads(int(n or os.cpu_count() or 1))
    return num_threads()


def spgemm_rowcounts(A, B, nB, row0=0, row1=None) -> np.ndarray:
    """nnz of every row of C = A*B, rows [row0,row1): the count pass of the SPA only (no C is built)."""
    rpA, ciA = np.ascontiguousarray(A[0], np.int64), np.ascontiguousarray(A[1], np.int32)
    rpB, ciB = np.ascontiguousarray(B[0], np.int64), np.ascontiguousarray(B[1], np.int32)
    mA = rpA.size - 1
    if row1 is None:
        row1 = mA
    out = np.zeros(max(row1 - row0, 0), np.int64)
    rc = lib().orc_spgemm_rowcounts(int(mA), int(nB), _p(rpA, C.c_int64), _p(ciA, C.c_int), _p(rpB, C.c_int64), _p(ciB, C.c_int),
                                    int(row0), int(row1), _p(out, C.c_int64))
    assert rc == 0
    return out
