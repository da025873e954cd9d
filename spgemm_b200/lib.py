"""ctypes binding of libtilespgemm_b200.so (C ABI: include/tilespgemm.h).

There is no Python or CPU implementation behind this module: if the shared library is missing the
import of the handle fails loudly, and if no CUDA device is usable every call raises TsgError.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtilespgemm_b200.so")

# every symbol include/tilespgemm.h declares (tests/test_abi.py checks the header against this list)
EXPORTS = [
    "csr2tile_row_major", "csr2tile_col_major", "tilespgemm", "tile2csr", "matrix_destroy", "matrix_transposition",
    "tilespgemm_last_error", "tilespgemm_last_error_string", "tilespgemm_clear_error",
    "tsg_init", "tsg_shutdown", "tsg_stream", "tsg_sync", "tsg_launch_count", "tsg_timer_start", "tsg_timer_stop",
    "tsg_csr_upload", "tsg_csr_wrap", "tsg_csr_download", "tsg_csr_free", "tsg_csr_validate", "tsg_csr_row_slice",
    "tsg_csr_canonicalize",
    "tsg_transpose", "tsg_nnzcub", "tsg_csr2tile", "tsg_tile_upload", "tsg_tile_download", "tsg_tile_alloc",
    "tsg_tile_free", "tsg_tilerow_weights", "tsg_spgemm", "tsg_tile2csr", "tsg_tile_rowsums", "tsg_spgemm_csr_host",
    "tsg_spgemm_to_host", "tsg_spgemm_csr_host_into", "tsg_plan_slabs", "tsg_spgemm_slabs",
    # Part 3: general tile sizes
    "tsg_gtile_size_ok", "tsg_gtile_csr2tile", "tsg_gtile_upload", "tsg_gtile_download", "tsg_gtile_free", "tsg_gtile_spgemm",
    "tsg_gtile_tile2csr",
]


class SMatrix(C.Structure):
    """Mirror of the reference struct (src/common.h:150-172) / include/tilespgemm.h."""
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


class DCsr(C.Structure):
    _fields_ = [("m", C.c_int), ("n", C.c_int), ("nnz", C.c_longlong),
                ("rowptr", C.c_void_p), ("colidx", C.c_void_p), ("val", C.c_void_p), ("owner", C.c_void_p)]


class DTile(C.Structure):
    _fields_ = [
        ("m", C.c_int), ("n", C.c_int), ("tilem", C.c_int), ("tilen", C.c_int), ("numtile", C.c_int),
        ("col_major", C.c_int), ("trow0", C.c_int), ("cached", C.c_int), ("nnz", C.c_longlong),
        ("tile_ptr", C.c_void_p), ("tile_columnidx", C.c_void_p), ("tile_rowidx", C.c_void_p), ("tile_nnz", C.c_void_p),
        ("val", C.c_void_p), ("col", C.c_void_p), ("ptr", C.c_void_p), ("mask", C.c_void_p),
        ("csc_tile_ptr", C.c_void_p), ("csc_tile_rowidx", C.c_void_p), ("rm2csc", C.c_void_p),
        ("pat", C.c_void_p), ("npat", C.c_int),
        ("slab", C.c_void_p * 4), ("slab_bytes", C.c_size_t * 4),
    ]


class GTile(C.Structure):
    """tsg_gtile: a tiled matrix on the device with tiles of tile_rows x tile_cols (include/tilespgemm.h, Part 3)."""
    _fields_ = [
        ("m", C.c_int), ("n", C.c_int), ("tile_rows", C.c_int), ("tile_cols", C.c_int),
        ("tilem", C.c_int), ("tilen", C.c_int), ("numtile", C.c_int), ("col_major", C.c_int), ("nnz", C.c_longlong),
        ("tile_ptr", C.c_void_p), ("tile_columnidx", C.c_void_p), ("tile_rowidx", C.c_void_p), ("tile_nnz", C.c_void_p),
        ("val", C.c_void_p), ("col", C.c_void_p), ("ptr", C.c_void_p), ("mask", C.c_void_p),
        ("csc_tile_ptr", C.c_void_p), ("csc_tile_rowidx", C.c_void_p), ("rm2csc", C.c_void_p),
        ("slab", C.c_void_p * 2), ("slab_bytes", C.c_size_t * 2),
    ]


class Stats(C.Structure):
    _fields_ = [("numblkC", C.c_longlong), ("nnzC", C.c_longlong), ("pairs", C.c_longlong),
                ("ms_step1", C.c_double), ("ms_step2", C.c_double), ("ms_step3", C.c_double),
                ("ms_alloc", C.c_double), ("ms_total", C.c_double),
                ("algorithmic_bytes", C.c_longlong), ("launches", C.c_int),
                ("rows_staged", C.c_int), ("rows_gather", C.c_int), ("tiles_dense", C.c_int), ("rows_smem", C.c_int), ("tiles_nonempty", C.c_int), ("plan_recipes", C.c_int), ("row_templates", C.c_int)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class TsgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"tilespgemm_b200 error {code}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """Load the CUDA library. Raises if it was not built -- there is nothing to fall back to."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). spgemm_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.tilespgemm_last_error_string.restype = C.c_char_p
    lib.tsg_stream.restype = C.c_void_p
    lib.tsg_launch_count.restype = C.c_longlong
    for name in ("csr2tile_row_major", "csr2tile_col_major", "tile2csr", "matrix_destroy", "tilespgemm",
                 "matrix_transposition", "tilespgemm_clear_error", "tsg_shutdown", "tsg_csr_free", "tsg_tile_free", "tsg_gtile_free"):
        getattr(lib, name).restype = None
    _lib = lib
    return lib


def check(rc: int = 0) -> None:
    """Raise the library's latched error (drop-in entry points are void and only latch)."""
    lib = load()
    code = rc or lib.tilespgemm_last_error()
    if code:
        msg = lib.tilespgemm_last_error_string().decode()
        lib.tilespgemm_clear_error()
        raise TsgError(code, msg)
