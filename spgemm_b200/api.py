"""Host-side mirror of the reference driver's interface for the TileSpGEMM hot path.

Two layers over the C ABI (include/tilespgemm.h), both thin:

* the reference-named calls -- `csr2tile_row_major`, `csr2tile_col_major`, `tilespgemm`, `tile2csr`,
  `matrix_destroy`, `matrix_transposition` -- operating on a `HostMatrix` that wraps the same
  `SMatrix` struct the reference driver allocates (reference src/main.cu:77-152, 168, 191, 261, 327);
  host buffers in, host buffers out, like the reference;
* the device-resident calls (`DeviceCSR`, `DeviceTiled`, `csr2tile`, `spgemm`, `tile2csr_device`, ...)
  that keep everything in HBM between the stages and run C tile-row slabs.

Everything computes on the GPU through libtilespgemm_b200.so; errors raise `TsgError`.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as L
from .lib import Stats, TsgError  # noqa: F401  (re-exported)

_libc = C.CDLL(None)
_libc.free.argtypes = [C.c_void_p]
_libc.free.restype = None


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def _take(ptr, n, dtype):
    """Copy n items out of a malloc()ed C array into numpy."""
    if n <= 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


# ----------------------------------------------------------------------------------------------
# Reference-named layer
# ----------------------------------------------------------------------------------------------
class HostMatrix:
    """An `SMatrix` (reference src/common.h:150-172) plus the numpy arrays backing its CSR view."""

    def __init__(self):
        self.s = L.SMatrix()
        self._keep = []          # numpy arrays whose memory the struct points into
        self._tile_owned = False  # tile arrays malloc()ed by the library
        self._csr_owned = False   # CSR arrays malloc()ed by the library (tile2csr)

    @classmethod
    def from_csr(cls, m, n, rowptr, colidx, val):
        self = cls()
        rp = np.ascontiguousarray(rowptr, dtype=np.int32)
        ci = np.ascontiguousarray(colidx, dtype=np.int32)
        v = np.ascontiguousarray(val, dtype=np.float64)
        self._keep = [rp, ci, v]
        s = self.s
        s.m, s.n, s.nnz, s.isSymmetric = int(m), int(n), int(rp[m]), 0
        s.rowpointer, s.columnindex, s.value = _p(rp, C.c_int), _p(ci, C.c_int), _p(v, C.c_double)
        return self

    def alias_csr_of(self, other: "HostMatrix"):
        """B aliases A's CSR arrays for -aat 0 (reference src/main.cu:145-151)."""
        s, o = self.s, other.s
        s.m, s.n, s.nnz, s.isSymmetric = o.m, o.n, o.nnz, 0
        s.rowpointer, s.columnindex, s.value = o.rowpointer, o.columnindex, o.value
        self._keep = other._keep
        return self

    # numpy copies of the library-produced arrays. tile_rows x tile_cols = the size of ONE tile of this matrix
    # (general tile sizes: A tile_size_m x tile_size_n, B tile_size_n x tile_size_m, C tile_size_m x tile_size_m)
    def tiles(self, tile_rows: int = 16, tile_cols: int = 16) -> dict:
        s = self.s
        nt, nnz = s.numtile, s.nnz
        d = dict(m=s.m, n=s.n, tilem=s.tilem, tilen=s.tilen, numtile=nt, nnz=nnz,
                 tile_ptr=_take(s.tile_ptr, s.tilem + 1, np.int32),
                 tile_columnidx=_take(s.tile_columnidx, nt, np.int32),
                 tile_rowidx=_take(s.tile_rowidx, nt, np.int32),
                 tile_nnz=_take(s.tile_nnz, nt + 1, np.int32),
                 val=_take(s.tile_csr_Value, nnz, np.float64),
                 col=_take(s.tile_csr_Col, nnz, np.uint16),
                 ptr=_take(s.tile_csr_Ptr, nt * tile_rows, np.uint16),
                 mask=_take(s.mask, nt * tile_rows * (tile_cols // 16), np.uint16))
        if s.csc_tile_ptr:
            d["csc_tile_ptr"] = _take(s.csc_tile_ptr, s.tilen + 1, np.int32)
            d["csc_tile_rowidx"] = _take(s.csc_tile_rowidx, nt, np.int32)
        return d

    def csr(self):
        s = self.s
        return (_take(s.rowpointer, s.m + 1, np.int32), _take(s.columnindex, s.nnz, np.int32),
                _take(s.value, s.nnz, np.float64))


def csr2tile_row_major(matrix: HostMatrix, tile_size_m: int = 16, tile_size_n: int = 16) -> None:
    """Reference src/csr2tile.h:205."""
    L.load().csr2tile_row_major(C.byref(matrix.s), tile_size_m, tile_size_n)
    L.check()
    matrix._tile_owned = True


def csr2tile_col_major(matrix: HostMatrix, tile_size_m: int = 16, tile_size_n: int = 16) -> None:
    """Reference src/csr2tile.h:279."""
    L.load().csr2tile_col_major(C.byref(matrix.s), tile_size_m, tile_size_n)
    L.check()
    matrix._tile_owned = True


def tilespgemm(A: HostMatrix, B: HostMatrix, nnzCub: int, tile_size_m: int = 16, tile_size_n: int = 16):
    """Reference src/tilespgemm-cuda.h:2220. Returns (C, info) with the reference's out-parameters."""
    Cm = HostMatrix()
    nnzC = C.c_ulonglong(0)
    outs = [C.c_double(0) for _ in range(7)]  # compression_rate, time_tile, gflops_tile, step1, step2, step3, malloc
    L.load().tilespgemm(C.byref(A.s), C.byref(B.s), C.byref(Cm.s), None, None, 0, C.c_double(0), C.c_double(0),
                        C.c_ulonglong(int(nnzCub)), C.byref(nnzC), C.byref(outs[0]), C.byref(outs[1]), C.byref(outs[2]),
                        b"synthetic", C.byref(outs[3]), C.byref(outs[4]), C.byref(outs[5]), C.byref(outs[6]),
                        tile_size_m, tile_size_n)
    L.check()
    Cm._tile_owned = True
    info = dict(nnzC_computed=nnzC.value, compression_rate=outs[0].value, time_tile=outs[1].value,
                gflops_tile=outs[2].value, time_step1=outs[3].value, time_step2=outs[4].value,
                time_step3=outs[5].value, time_malloc=outs[6].value)
    return Cm, info


def tile2csr(matrix: HostMatrix, tile_size_m: int = 16, tile_size_n: int = 16) -> None:
    """Reference src/tile2csr.h:72 (the driver passes (tile_size_m, tile_size_m), main.cu:327)."""
    L.load().tile2csr(C.byref(matrix.s), tile_size_m, tile_size_n)
    L.check()
    matrix._csr_owned = True


def matrix_destroy(matrix: HostMatrix) -> None:
    """Reference src/csr2tile.h:509, plus the arrays the reference leaks."""
    s = matrix.s
    if matrix._tile_owned:
        L.load().matrix_destroy(C.byref(s))
        for name in ("tile_rowidx", "csc_tile_ptr", "csc_tile_rowidx"):
            ptr = getattr(s, name)
            if ptr:
                _libc.free(C.cast(ptr, C.c_void_p))
            setattr(s, name, None)
        matrix._tile_owned = False
    if matrix._csr_owned:
        for name in ("rowpointer", "columnindex", "value"):
            ptr = getattr(s, name)
            if ptr:
                _libc.free(C.cast(ptr, C.c_void_p))
            setattr(s, name, None)
        matrix._csr_owned = False


def matrix_transposition(m, n, rowptr, colidx, val):
    """Reference src/utils.h:161. Returns (cscColPtr, cscRowIdx, cscVal)."""
    rp = np.ascontiguousarray(rowptr, dtype=np.int32)
    ci = np.ascontiguousarray(colidx, dtype=np.int32)
    v = np.ascontiguousarray(val, dtype=np.float64)
    nnz = int(rp[m])
    colptr = np.zeros(n + 1, np.int32)
    rowidx = np.zeros(max(nnz, 1), np.int32)
    cv = np.zeros(max(nnz, 1), np.float64)
    L.load().matrix_transposition(int(m), int(n), nnz, _p(rp, C.c_int), _p(ci, C.c_int), _p(v, C.c_double),
                                  _p(rowidx, C.c_int), _p(colptr, C.c_int), _p(cv, C.c_double))
    L.check()
    return colptr, rowidx[:nnz], cv[:nnz]


# ----------------------------------------------------------------------------------------------
# Device-resident layer
# ----------------------------------------------------------------------------------------------
def init(device: int = 0) -> None:
    L.check(L.load().tsg_init(int(device)))


def shutdown() -> None:
    L.load().tsg_shutdown()


def launch_count() -> int:
    return int(L.load().tsg_launch_count())


def sync() -> None:
    L.check(L.load().tsg_sync())


def timer_start() -> None:
    L.check(L.load().tsg_timer_start())


def timer_stop() -> float:
    """Milliseconds the library stream spent since timer_start() (CUDA events, host gaps included)."""
    ms = C.c_double(0)
    L.check(L.load().tsg_timer_stop(C.byref(ms)))
    return ms.value


class DeviceCSR:
    def __init__(self):
        self.d = L.DCsr()

    @classmethod
    def upload(cls, m, n, rowptr, colidx, val):
        self = cls()
        rp = np.ascontiguousarray(rowptr, dtype=np.int32)
        ci = np.ascontiguousarray(colidx, dtype=np.int32)
        v = np.ascontiguousarray(val, dtype=np.float64)
        L.check(L.load().tsg_csr_upload(int(m), int(n), _p(rp, C.c_int), _p(ci, C.c_int), _p(v, C.c_double), C.byref(self.d)))
        return self

    @classmethod
    def upload_ptrs(cls, m, n, rowptr_addr, colidx_addr, val_addr):
        """Upload from raw host addresses (e.g. pinned torch tensors' data_ptr())."""
        self = cls()
        L.check(L.load().tsg_csr_upload(int(m), int(n), C.c_void_p(rowptr_addr), C.c_void_p(colidx_addr),
                                        C.c_void_p(val_addr), C.byref(self.d)))
        return self

    @classmethod
    def wrap(cls, m, n, nnz, d_rowptr, d_colidx, d_val):
        self = cls()
        L.check(L.load().tsg_csr_wrap(int(m), int(n), C.c_longlong(int(nnz)), C.c_void_p(d_rowptr), C.c_void_p(d_colidx),
                                      C.c_void_p(d_val), C.byref(self.d)))
        return self

    m = property(lambda self: self.d.m)
    n = property(lambda self: self.d.n)
    nnz = property(lambda self: self.d.nnz)

    def download(self):
        rp = np.zeros(self.d.m + 1, np.int32)
        ci = np.zeros(max(self.d.nnz, 1), np.int32)
        v = np.zeros(max(self.d.nnz, 1), np.float64)
        L.check(L.load().tsg_csr_download(C.byref(self.d), _p(rp, C.c_int), _p(ci, C.c_int), _p(v, C.c_double)))
        return rp, ci[:self.d.nnz], v[:self.d.nnz]

    def download_into(self, rowptr_addr, colidx_addr, val_addr):
        L.check(L.load().tsg_csr_download(C.byref(self.d), C.c_void_p(rowptr_addr), C.c_void_p(colidx_addr),
                                          C.c_void_p(val_addr)))

    def row_slice(self, row0: int, row1: int) -> "DeviceCSR":
        """Rows [row0, row1) as a CSR of its own (rebased row pointer; colidx / val borrowed from self, which must
        outlive the slice)."""
        out = DeviceCSR()
        L.check(L.load().tsg_csr_row_slice(C.byref(self.d), int(row0), int(row1), C.byref(out.d)))
        out._parent = self
        return out

    def canonicalize(self, dup_policy: str = "sum") -> "DeviceCSR":
        """Sorted, duplicate-free copy (tsg_csr_canonicalize); dup_policy 'sum' or 'first'."""
        out = DeviceCSR()
        L.check(L.load().tsg_csr_canonicalize(C.byref(self.d), {"sum": 0, "first": 1}[dup_policy], C.byref(out.d)))
        return out

    def free(self):
        L.load().tsg_csr_free(C.byref(self.d))


class DeviceTiled:
    def __init__(self):
        self.d = L.DTile()

    def __getattr__(self, k):
        if k in ("m", "n", "tilem", "tilen", "numtile", "nnz", "col_major", "trow0"):
            return getattr(self.d, k)
        raise AttributeError(k)

    def download(self) -> dict:
        h = HostMatrix()
        L.check(L.load().tsg_tile_download(C.byref(self.d), C.byref(h.s)))
        h._tile_owned = True
        out = h.tiles()
        out["trow0"] = self.d.trow0
        matrix_destroy(h)
        return out

    def slab_bytes(self):
        return [int(b) for b in self.d.slab_bytes]

    def free(self):
        L.load().tsg_tile_free(C.byref(self.d))


def csr2tile(a: DeviceCSR, col_major: bool) -> DeviceTiled:
    t = DeviceTiled()
    L.check(L.load().tsg_csr2tile(C.byref(a.d), int(bool(col_major)), C.byref(t.d)))
    return t


def tile_alloc(m, n, numtile, nnz, col_major) -> DeviceTiled:
    t = DeviceTiled()
    L.check(L.load().tsg_tile_alloc(int(m), int(n), int(numtile), C.c_longlong(int(nnz)), int(bool(col_major)), C.byref(t.d)))
    return t


def tile_upload(h: HostMatrix, col_major: bool) -> DeviceTiled:
    t = DeviceTiled()
    L.check(L.load().tsg_tile_upload(C.byref(h.s), int(bool(col_major)), C.byref(t.d)))
    return t


# ----------------------------------------------------------------------------------------------
# General tile sizes (include/tilespgemm.h Part 3, csrc/gentile.cu)
# ----------------------------------------------------------------------------------------------
class DeviceGTiled:
    """A tiled matrix on the device whose tiles are tile_rows x tile_cols (multiples of 16 up to 128)."""

    def __init__(self):
        self.d = L.GTile()

    def __getattr__(self, k):
        if k in ("m", "n", "tilem", "tilen", "numtile", "nnz", "col_major", "tile_rows", "tile_cols"):
            return getattr(self.d, k)
        raise AttributeError(k)

    def download(self) -> dict:
        h = HostMatrix()
        L.check(L.load().tsg_gtile_download(C.byref(self.d), C.byref(h.s)))
        h._tile_owned = True
        out = h.tiles(self.d.tile_rows, self.d.tile_cols)
        matrix_destroy(h)
        return out

    def free(self):
        L.load().tsg_gtile_free(C.byref(self.d))


def gtile_size_ok(tile_rows: int, tile_cols: int) -> bool:
    return bool(L.load().tsg_gtile_size_ok(int(tile_rows), int(tile_cols)))


def gtile_csr2tile(a: DeviceCSR, col_major: bool, tile_rows: int, tile_cols: int) -> DeviceGTiled:
    """CSR -> tiles of tile_rows x tile_cols. For the reference's csr2tile_col_major(B, tile_size_m, tile_size_n) pass
    tile_rows = tile_size_n, tile_cols = tile_size_m."""
    t = DeviceGTiled()
    L.check(L.load().tsg_gtile_csr2tile(C.byref(a.d), int(bool(col_major)), int(tile_rows), int(tile_cols), C.byref(t.d)))
    return t


def gtile_upload(h: HostMatrix, col_major: bool, tile_rows: int, tile_cols: int) -> DeviceGTiled:
    t = DeviceGTiled()
    L.check(L.load().tsg_gtile_upload(C.byref(h.s), int(bool(col_major)), int(tile_rows), int(tile_cols), C.byref(t.d)))
    return t


def gtile_spgemm(a: DeviceGTiled, b: DeviceGTiled):
    """Steps 1-3 for general tiles. Returns (C, stats dict)."""
    c = DeviceGTiled()
    st = L.Stats()
    L.check(L.load().tsg_gtile_spgemm(C.byref(a.d), C.byref(b.d), C.byref(c.d), C.byref(st)))
    return c, st.as_dict()


def gtile_tile2csr(t: DeviceGTiled) -> DeviceCSR:
    out = DeviceCSR()
    L.check(L.load().tsg_gtile_tile2csr(C.byref(t.d), C.byref(out.d)))
    return out


def transpose(a: DeviceCSR) -> DeviceCSR:
    at = DeviceCSR()
    L.check(L.load().tsg_transpose(C.byref(a.d), C.byref(at.d)))
    return at


def nnzcub(a: DeviceCSR, b: DeviceCSR) -> int:
    out = C.c_ulonglong(0)
    L.check(L.load().tsg_nnzcub(C.byref(a.d), C.byref(b.d), C.byref(out)))
    return int(out.value)


def tilerow_weights(a: DeviceTiled, b: DeviceTiled) -> np.ndarray:
    w = np.zeros(max(a.tilem, 1), np.int64)
    L.check(L.load().tsg_tilerow_weights(C.byref(a.d), C.byref(b.d), _p(w, C.c_longlong)))
    return w[:a.tilem]


def spgemm(a: DeviceTiled, b: DeviceTiled, trow0: int = 0, trow1: int = -1):
    """Steps 1-3 for C tile-rows [trow0, trow1). Returns (C slab, stats dict)."""
    c = DeviceTiled()
    st = L.Stats()
    L.check(L.load().tsg_spgemm(C.byref(a.d), C.byref(b.d), int(trow0), int(trow1), C.byref(c.d), C.byref(st)))
    return c, st.as_dict()


def tile2csr_device(t: DeviceTiled) -> DeviceCSR:
    out = DeviceCSR()
    L.check(L.load().tsg_tile2csr(C.byref(t.d), C.byref(out.d)))
    return out


def tile_rowsums(t: DeviceTiled):
    """Per-row sums (C * ones) and per-row entry counts of a row-major tiled matrix / C slab."""
    s = np.zeros(max(t.m, 1), np.float64)
    c = np.zeros(max(t.m, 1), np.int64)
    L.check(L.load().tsg_tile_rowsums(C.byref(t.d), _p(s, C.c_double), _p(c, C.c_longlong)))
    return s[:t.m], c[:t.m]


def plan_slabs(weights: np.ndarray, max_pairs: int):
    """Cut tile-rows into contiguous slabs of at most `max_pairs` matched tile pairs each (a single
    tile-row heavier than that gets a slab of its own)."""
    w = np.asarray(weights, dtype=np.int64)
    cuts, acc = [0], 0
    for i, x in enumerate(w):
        if acc > 0 and acc + x > max_pairs:
            cuts.append(i)
            acc = 0
        acc += int(x)
    cuts.append(len(w))
    return [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]


def plan_to_host_slabs(weights: np.ndarray, nslabs: int = 0, trow0: int = 0, trow1: int = -1):
    """The slab boundaries tsg_spgemm_to_host would use for these step-1 weights (host-only: tsg_plan_slabs)."""
    w = np.ascontiguousarray(weights, dtype=np.int64)
    if trow1 < 0:
        trow1 = len(w)
    cuts = np.zeros(max(trow1 - trow0, 0) + 2, np.int32)
    n = L.load().tsg_plan_slabs(_p(w, C.c_longlong), int(trow0), int(trow1), int(nslabs), _p(cuts, C.c_int), int(cuts.size))
    if n < 0:
        L.check()
    return [int(x) for x in cuts[:n + 1]] if n > 0 else []


def spgemm_slabs(a: DeviceTiled, b: DeviceTiled, max_pairs: int = 1 << 28, sink=None, trow0: int = 0, trow1: int = -1,
                 weights: np.ndarray | None = None):
    """Steps 1-3 over tile-rows [trow0, trow1) executed slab by slab with a bounded device footprint
    (R-MAT products do not fit one GPU whole, nor int32 offsets). Each slab is handed
    to `sink(C_slab, stats)` and freed. Returns the totals (64-bit) and the per-slab stats."""
    if trow1 < 0:
        trow1 = a.tilem
    w = tilerow_weights(a, b) if weights is None else weights
    slabs = [(t0 + trow0, t1 + trow0) for t0, t1 in plan_slabs(w[trow0:trow1], max_pairs)]
    tot = dict(numblkC=0, nnzC=0, pairs=0, ms_step1=0.0, ms_step2=0.0, ms_step3=0.0, ms_alloc=0.0, ms_total=0.0,
               algorithmic_bytes=0, launches=0, rows_staged=0, rows_gather=0, tiles_dense=0, rows_smem=0, tiles_nonempty=0, plan_recipes=0, row_templates=0, slabs=len(slabs))
    per = []
    for t0, t1 in slabs:
        c, st = spgemm(a, b, t0, t1)
        if sink is not None:
            sink(c, st)
        c.free()
        for k in tot:
            if k in ("rows_smem", "plan_recipes", "row_templates"):
                tot[k] = max(tot[k], st[k])
            elif k != "slabs":
                tot[k] += st[k]
        per.append(dict(st, trow0=t0, trow1=t1))
    return tot, per


def spgemm_to_host(a: DeviceTiled, b: DeviceTiled, rowptr_addr: int, colidx_addr: int, val_addr: int, cap: int,
                   nslabs: int = 0, trow0: int = 0, trow1: int = -1):
    """Steps 1-3 + tile2csr + D2H for C tile-rows [trow0, trow1), slab by slab with the copy of one slab
    overlapping the computation of the next (include/tilespgemm.h: tsg_spgemm_to_host). The three addresses
    are HOST buffers (pinned for the overlap): rows+1 int32, `cap` int32, `cap` float64. Returns
    (nnz, stats dict); raises TsgError (code TSG_ERR_NOMEM, message naming the size needed) if cap is too small."""
    nnz = C.c_longlong(0)
    st = L.Stats()
    L.check(L.load().tsg_spgemm_to_host(C.byref(a.d), C.byref(b.d), int(trow0), int(trow1), int(nslabs), C.c_void_p(rowptr_addr),
                                        C.c_void_p(colidx_addr), C.c_void_p(val_addr), C.c_longlong(int(cap)), C.byref(nnz),
                                        C.byref(st)))
    return nnz.value, st.as_dict()


def spgemm_csr_host_into(m, k, n, A, out, B=None, aat=False):
    """Host CSR in -> host CSR out into caller arrays, overlapped (tsg_spgemm_csr_host_into). A, B = (rowptr, colidx, val);
    out = (rowptr int32[m+1], colidx int32[cap], val float64[cap]) numpy arrays (pinned memory for full PCIe speed).
    Returns (nnz, stats dict)."""
    rpA, ciA, vA = (np.ascontiguousarray(A[0], np.int32), np.ascontiguousarray(A[1], np.int32),
                    np.ascontiguousarray(A[2], np.float64))
    if B is not None:
        rpB, ciB, vB = (np.ascontiguousarray(B[0], np.int32), np.ascontiguousarray(B[1], np.int32),
                        np.ascontiguousarray(B[2], np.float64))
        bargs = (_p(rpB, C.c_int), _p(ciB, C.c_int), _p(vB, C.c_double))
    else:
        bargs = (None, None, None)
    orp, oci, ov = out
    assert orp.dtype == np.int32 and oci.dtype == np.int32 and ov.dtype == np.float64 and orp.size >= m + 1
    cap = min(oci.size, ov.size)
    nnz = C.c_longlong(0)
    st = L.Stats()
    L.check(L.load().tsg_spgemm_csr_host_into(int(m), int(k), int(n), _p(rpA, C.c_int), _p(ciA, C.c_int), _p(vA, C.c_double),
                                              *bargs, int(bool(aat)), _p(orp, C.c_int), _p(oci, C.c_int), _p(ov, C.c_double),
                                              C.c_longlong(int(cap)), C.byref(nnz), C.byref(st)))
    return nnz.value, st.as_dict()


def spgemm_csr_host(m, k, n, A, B=None, aat=False):
    """Whole pipeline, host CSR in -> host CSR out. A, B = (rowptr, colidx, val); B=None means B=A
    (or A^T with aat=True). Returns (rowptr, colidx, val, stats dict)."""
    rpA, ciA, vA = (np.ascontiguousarray(A[0], np.int32), np.ascontiguousarray(A[1], np.int32),
                    np.ascontiguousarray(A[2], np.float64))
    if B is not None:
        rpB, ciB, vB = (np.ascontiguousarray(B[0], np.int32), np.ascontiguousarray(B[1], np.int32),
                        np.ascontiguousarray(B[2], np.float64))
        bargs = (_p(rpB, C.c_int), _p(ciB, C.c_int), _p(vB, C.c_double))
    else:
        bargs = (None, None, None)
    crp, cci, cv = C.POINTER(C.c_int)(), C.POINTER(C.c_int)(), C.POINTER(C.c_double)()
    cnnz = C.c_longlong(0)
    st = L.Stats()
    rc = L.load().tsg_spgemm_csr_host(int(m), int(k), int(n), _p(rpA, C.c_int), _p(ciA, C.c_int), _p(vA, C.c_double),
                                      *bargs, int(bool(aat)), C.byref(crp), C.byref(cci), C.byref(cv), C.byref(cnnz),
                                      C.byref(st))
    L.check(rc)
    out = (_take(crp, m + 1, np.int32), _take(cci, cnnz.value, np.int32), _take(cv, cnnz.value, np.float64), st.as_dict())
    for ptr in (crp, cci, cv):
        _libc.free(C.cast(ptr, C.c_void_p))
    return out
