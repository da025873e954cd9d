"""ctypes front-end of the reference's OWN CPU functions (oracle/_ref/*.so, built in place from
/root/reference/src by oracle/Makefile; see ref_shim.cpp).

TEST INFRASTRUCTURE ONLY: validates the restatement (oracle/spa_ref.c) and generates the golden
vectors under tests/golden/. `available()` is False when oracle/_ref was never built (e.g. a box
without /root/reference and without the prebuilt files); callers must skip, not fall back.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .oracle import Tiled

_HERE = os.path.dirname(os.path.abspath(__file__))
_CPU = os.path.join(_HERE, "_ref", "libref_cpu.so")
_SPA = os.path.join(_HERE, "_ref", "libref_spa.so")
_lib_cpu = None
_lib_spa = None


class SMatrix(C.Structure):
    """Field-for-field mirror of the reference struct (src/common.h:150-172)."""
    _fields_ = [
        ("m", C.c_int), ("n", C.c_int), ("nnz", C.c_int), ("isSymmetric", C.c_int),
        ("value", C.POINTER(C.c_double)), ("columnindex", C.POINTER(C.c_int)), ("rowpointer", C.POINTER(C.c_int)),
        ("tilem", C.c_int), ("tilen", C.c_int),
        ("tile_ptr", C.POINTER(C.c_int)), ("tile_columnidx", C.POINTER(C.c_int)),
        ("tile_rowidx", C.POINTER(C.c_int)), ("tile_nnz", C.POINTER(C.c_int)),
        ("numtile", C.c_int),
        ("tile_csr_Value", C.POINTER(C.c_double)), ("tile_csr_Col", C.POINTER(C.c_uint16)),
        ("tile_csr_Ptr", C.POINTER(C.c_uint16)), ("mask", C.POINTER(C.c_uint16)),
        ("csc_tile_ptr", C.POINTER(C.c_int)), ("csc_tile_rowidx", C.POINTER(C.c_int)),
    ]


def available() -> bool:
    return os.path.exists(_CPU) and os.path.exists(_SPA)


def _cpu():
    global _lib_cpu
    if _lib_cpu is None:
        _lib_cpu = C.CDLL(_CPU)
        assert _lib_cpu.ref_sizeof_smatrix() == C.sizeof(SMatrix), "SMatrix mirror out of sync with src/common.h"
    return _lib_cpu


def _spa():
    global _lib_spa
    if _lib_spa is None:
        _lib_spa = C.CDLL(_SPA)
    return _lib_spa


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def _arr(p, n, dtype):
    if n <= 0 or not p:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(p, shape=(n,)).astype(dtype, copy=True)


def _csr2tile(m, n, rowptr, colidx, val, colmajor: bool, tile_size_m: int = 16, tile_size_n: int = 16) -> Tiled:
    rp = np.ascontiguousarray(rowptr, dtype=np.int32)
    ci = np.ascontiguousarray(colidx, dtype=np.int32)
    v = np.ascontiguousarray(val, dtype=np.float64)
    s = SMatrix()
    s.m, s.n, s.nnz, s.isSymmetric = int(m), int(n), int(rp[m]), 0
    s.rowpointer, s.columnindex, s.value = _p(rp, C.c_int), _p(ci, C.c_int), _p(v, C.c_double)
    if colmajor:
        _cpu().ref_csr2tile_col_major_g(C.byref(s), int(tile_size_m), int(tile_size_n))
    else:
        _cpu().ref_csr2tile_row_major_g(C.byref(s), int(tile_size_m), int(tile_size_n))
    nt, nnz = s.numtile, s.nnz
    # rows x columns of one tile: A's are tile_size_m x tile_size_n, B's tile_size_n x tile_size_m (src/main.cu:84)
    tr, tc = (tile_size_n, tile_size_m) if colmajor else (tile_size_m, tile_size_n)
    out = Tiled(
        m=s.m, n=s.n, tilem=s.tilem, tilen=s.tilen, numtile=nt, nnz=nnz,
        tile_ptr=_arr(s.tile_ptr, s.tilem + 1, np.int32),
        tile_columnidx=_arr(s.tile_columnidx, nt, np.int32),
        tile_rowidx=_arr(s.tile_rowidx, nt, np.int32),
        tile_nnz=_arr(s.tile_nnz, nt + 1, np.int64),
        val=_arr(s.tile_csr_Value, nnz, np.float64),
        col=_arr(s.tile_csr_Col, nnz, np.uint16),
        ptr=_arr(s.tile_csr_Ptr, nt * tr, np.uint16),
        mask=_arr(s.mask, nt * tr * (tc // 16), np.uint16),
        csc_tile_ptr=_arr(s.csc_tile_ptr, s.tilen + 1, np.int32) if colmajor else None,
        csc_tile_rowidx=_arr(s.csc_tile_rowidx, nt, np.int32) if colmajor else None,
        tr=tr, tc=tc,
    )
    _cpu().ref_matrix_destroy(C.byref(s))  # frees what src/csr2tile.h:509 frees (the rest leaks, as upstream)
    return out


def csr2tile_row_major(m, n, rowptr, colidx, val, tile_size_m=16, tile_size_n=16) -> Tiled:
    """Reference src/csr2tile.h:205, unmodified."""
    return _csr2tile(m, n, rowptr, colidx, val, False, tile_size_m, tile_size_n)


def csr2tile_col_major(m, n, rowptr, colidx, val, tile_size_m=16, tile_size_n=16) -> Tiled:
    """Reference src/csr2tile.h:279, unmodified (tiles of tile_size_n rows x tile_size_m columns)."""
    return _csr2tile(m, n, rowptr, colidx, val, True, tile_size_m, tile_size_n)


def tile2csr(t: Tiled):
    """Reference src/tile2csr.h:72, unmodified. Returns (rowptr, colidx, val)."""
    keep = {}

    def put(arr, dt):
        a = np.ascontiguousarray(arr, dtype=dt)
        keep[id(a)] = a
        return a

    s = SMatrix()
    s.m, s.n, s.tilem, s.tilen, s.numtile, s.nnz = t.m, t.n, t.tilem, t.tilen, t.numtile, t.nnz
    s.tile_ptr = _p(put(t.tile_ptr, np.int32), C.c_int)
    s.tile_columnidx = _p(put(t.tile_columnidx, np.int32), C.c_int)
    s.tile_nnz = _p(put(t.tile_nnz, np.int32), C.c_int)
    s.tile_csr_Value = _p(put(t.val, np.float64), C.c_double)
    s.tile_csr_Col = _p(put(t.col, np.uint16), C.c_uint16)
    s.tile_csr_Ptr = _p(put(t.ptr, np.uint16), C.c_uint16)
    _cpu().ref_tile2csr_g(C.byref(s), int(t.tr), int(t.tc))
    nnz = s.nnz
    out = (_arr(s.rowpointer, s.m + 1, np.int32), _arr(s.columnindex, nnz, np.int32), _arr(s.value, nnz, np.float64))
    for p in (s.rowpointer, s.columnindex, s.value):
        _cpu().ref_free(p)
    return out


def matrix_transposition(m, n, rowptr, colidx, val):
    """Reference src/utils.h:161, unmodified. Returns (colptr, rowidx, cscval)."""
    rp = np.ascontiguousarray(rowptr, dtype=np.int32)
    ci = np.ascontiguousarray(colidx, dtype=np.int32)
    v = np.ascontiguousarray(val, dtype=np.float64)
    nnz = int(rp[m])
    colptr = np.zeros(n + 1, np.int32)
    rowidx = np.zeros(nnz, np.int32)
    cv = np.zeros(nnz, np.float64)
    _cpu().ref_matrix_transposition(int(m), int(n), nnz, _p(rp, C.c_int), _p(ci, C.c_int), _p(v, C.c_double),
                                    _p(rowidx, C.c_int), _p(colptr, C.c_int), _p(cv, C.c_double))
    return colptr, rowidx, cv


def _spgemm(fn, A, B, nB, with_values):
    rpA, ciA, vA = (np.ascontiguousarray(A[0], np.int32), np.ascontiguousarray(A[1], np.int32),
                    np.ascontiguousarray(A[2], np.float64))
    rpB, ciB, vB = (np.ascontiguousarray(B[0], np.int32), np.ascontiguousarray(B[1], np.int32),
                    np.ascontiguousarray(B[2], np.float64))
    mA, mB = rpA.size - 1, rpB.size - 1
    rpC = np.zeros(mA + 1, np.int32)
    nnzC = C.c_int(0)
    dummy_i = np.zeros(1, np.int32)
    dummy_d = np.zeros(1, np.float64)
    args = [_p(rpA, C.c_int), _p(ciA, C.c_int), _p(vA, C.c_double), mA, mB, int(rpA[mA]),
            _p(rpB, C.c_int), _p(ciB, C.c_int), _p(vB, C.c_double), mB, int(nB), int(rpB[mB])]
    fn(*args, _p(rpC, C.c_int), _p(dummy_i, C.c_int), _p(dummy_d, C.c_double), mA, int(nB), C.byref(nnzC), 1)
    ciC = np.zeros(max(nnzC.value, 1), np.int32)
    vC = np.zeros(max(nnzC.value, 1), np.float64)
    fn(*args, _p(rpC, C.c_int), _p(ciC, C.c_int), _p(vC, C.c_double), mA, int(nB), C.byref(nnzC), 0)
    return rpC, ciC[:nnzC.value], (vC[:nnzC.value] if with_values else None)


def spgemm_spa(A, B, nB):
    """Reference src/spgemm_serialref_spa_new.h:7 (structure only), two-pass protocol."""
    rp, ci, _ = _spgemm(_cpu().ref_spgemm_spa, A, B, nB, False)
    return rp, ci


def spgemm_serialref(A, B, nB):
    """Reference src/external/cusparse/spgemm_serialref_spa.h:33 (dense-row SPA with values)."""
    return _spgemm(_spa().ref_spgemm_serialref, A, B, nB, True)
