"""Synthetic input matrices for the BASELINE.json configs, plus a MatrixMarket reader.

All generators return CSR `(m, n, rowptr int32, colidx int32, val float64)` with rows sorted and
duplicate-free (the library's input contract). SuiteSparse is not available offline,
so these stand in for the reference driver's `.mtx` loader (reference src/main.cu:97).

Value conventions:
  * "mod10"  : value[k] = k % 10 in CSR order -- what the reference driver does right after loading
               (src/main.cu:111-112); every partial sum is an exact integer, so GPU and CPU must
               agree bit-for-bit.
  * "hash"   : value[k] = 1 + (hash(k) mod 2^20)/2^20, positive, for the 1e-12 tolerance check.
"""
from __future__ import annotations

import numpy as np


def set_values(nnz: int, kind: str = "mod10") -> np.ndarray:
    k = np.arange(nnz, dtype=np.uint64)
    if kind == "mod10":
        return (k % np.uint64(10)).astype(np.float64)
    if kind == "hash":
        x = k * np.uint64(0x9E3779B97F4A7C15) + np.uint64(1)
        x ^= x >> np.uint64(29)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(32)
        return 1.0 + (x & np.uint64((1 << 20) - 1)).astype(np.float64) / float(1 << 20)
    if kind == "ones":
        return np.ones(nnz, dtype=np.float64)
    raise ValueError(kind)


def _from_coo(m, n, rows, cols, values="mod10"):
    """Sort by (row, col), merge duplicates, build CSR."""
    key = rows.astype(np.int64) * np.int64(n) + cols.astype(np.int64)
    key = np.unique(key)
    rows = (key // n).astype(np.int32)
    cols = (key % n).astype(np.int32)
    rowptr = np.zeros(m + 1, dtype=np.int64)
    rowptr[1:] = np.bincount(rows, minlength=m)
    rowptr = np.cumsum(rowptr)
    assert rowptr[-1] < 2**31, "nnz exceeds int32"
    return m, n, rowptr.astype(np.int32), cols, set_values(cols.size, values)


def lap2d(nx: int, ny: int | None = None, values="mod10"):
    """2D 5-point Laplacian, natural ordering. Config 1: nx=ny=256 -> n=65 536, nnz=326 656."""
    ny = ny or nx
    n = nx * ny
    idx = np.arange(n, dtype=np.int64)
    x, y = idx % nx, idx // nx
    r, c = [idx], [idx]
    for dx, dy in ((-1, 0), (1, 0), (0, -1), (0, 1)):
        ok = (x + dx >= 0) & (x + dx < nx) & (y + dy >= 0) & (y + dy < ny)
        r.append(idx[ok])
        c.append(idx[ok] + dx + dy * nx)
    return _from_coo(n, n, np.concatenate(r), np.concatenate(c), values)


def stencil27(nx: int, ny: int | None = None, nz: int | None = None, values="mod10"):
    """3D 27-point stencil, natural ordering (x fastest). Config 2: 128^3 -> n=2 097 152, nnz=55 742 968."""
    ny = ny or nx
    nz = nz or nx
    n = nx * ny * nz
    idx = np.arange(n, dtype=np.int64)
    x, y, z = idx % nx, (idx // nx) % ny, idx // (nx * ny)
    counts = np.zeros(n, dtype=np.int64)
    offs = [(dx, dy, dz) for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]  # ascending column order
    oks = []
    for dx, dy, dz in offs:
        ok = (x + dx >= 0) & (x + dx < nx) & (y + dy >= 0) & (y + dy < ny) & (z + dz >= 0) & (z + dz < nz)
        oks.append(ok)
        counts += ok
    rowptr = np.zeros(n + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(counts)
    nnz = int(rowptr[-1])
    assert nnz < 2**31
    colidx = np.empty(nnz, dtype=np.int32)
    cursor = rowptr[:-1].copy()
    for (dx, dy, dz), ok in zip(offs, oks):
        colidx[cursor[ok]] = (idx[ok] + dx + dy * nx + dz * nx * ny).astype(np.int32)
        cursor[ok] += 1
    return n, n, rowptr.astype(np.int32), colidx, set_values(nnz, values)


def rmat(scale: int, edge_factor: int = 16, a=0.57, b=0.19, c=0.19, d=0.05, seed=1, values="mod10"):
    """R-MAT (Chakrabarti et al.) with 2^scale vertices and edge_factor*2^scale edges before
    duplicate merging. Defaults are the Graph500 skew; configs 3/5 state the (a,b,c,d) they use."""
    n = 1 << scale
    ne = edge_factor * n
    rng = np.random.default_rng(seed)
    rows = np.zeros(ne, dtype=np.int64)
    cols = np.zeros(ne, dtype=np.int64)
    ab, abc = a + b, a + b + c
    for bit in range(scale):
        u = rng.random(ne)
        rbit = u >= ab                      # quadrants c, d -> row bit set
        cbit = ((u >= a) & (u < ab)) | (u >= abc)  # quadrants b, d -> col bit set
        rows |= rbit.astype(np.int64) << bit
        cols |= cbit.astype(np.int64) << bit
    return _from_coo(n, n, rows, cols, values)


def blockfem(nodes: int, dof: int = 6, band: int = 1, values="mod10"):
    """Banded block-FEM: block (i,j) is a dense dof x dof block for |i-j| <= band.
    Config 4: nodes=333 334, dof=6, band=1 -> n=2 000 004, nnz=36 000 000."""
    n = nodes * dof
    bi = np.arange(nodes, dtype=np.int64)
    lo = np.maximum(bi - band, 0)
    hi = np.minimum(bi + band, nodes - 1)
    nb = hi - lo + 1                                  # blocks per block-row
    row_len = np.repeat(nb * dof, dof)                # entries per scalar row
    rowptr = np.zeros(n + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(row_len)
    nnz = int(rowptr[-1])
    assert nnz < 2**31
    row_of = np.repeat(np.arange(n, dtype=np.int64), row_len)
    k = np.arange(nnz, dtype=np.int64) - rowptr[row_of]
    colidx = (lo[row_of // dof] * dof + k).astype(np.int32)
    return n, n, rowptr.astype(np.int32), colidx, set_values(nnz, values)


def random_sparse(m: int, n: int, density: float, seed=0, values="mod10"):
    """Uniform random pattern (tests: ragged edges, rectangular shapes)."""
    rng = np.random.default_rng(seed)
    nnz = max(int(m * n * density), 1)
    rows = rng.integers(0, m, nnz)
    cols = rng.integers(0, n, nnz)
    return _from_coo(m, n, rows, cols, values)


def read_mtx(path: str, values=None):
    """Minimal MatrixMarket coordinate reader (real/integer/pattern, general/symmetric), with the
    sort + duplicate merge the reference's loader lacks (src/mmio_highlevel.h:593-759 does neither).
    `values=None` keeps the file's values (last duplicate wins); a string applies set_values()."""
    with open(path) as f:
        header = f.readline().lower().split()
        assert header[0] == "%%matrixmarket" and header[2] == "coordinate", "coordinate format only"
        field, symm = header[3], header[4]
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        m, n, nz = (int(x) for x in line.split())
        data = np.loadtxt(f, ndmin=2) if nz else np.zeros((0, 3))
    rows = data[:, 0].astype(np.int64) - 1
    cols = data[:, 1].astype(np.int64) - 1
    vals = data[:, 2] if field != "pattern" and data.shape[1] > 2 else np.ones(rows.size)
    if symm in ("symmetric", "skew-symmetric", "hermitian"):
        off = rows != cols
        rows, cols = np.concatenate([rows, cols[off]]), np.concatenate([cols, rows[off]])
        vals = np.concatenate([vals, vals[off] * (-1.0 if symm == "skew-symmetric" else 1.0)])
    key = rows * n + cols
    order = np.argsort(key, kind="stable")
    key, vals = key[order], vals[order]
    last = np.ones(key.size, dtype=bool)
    last[:-1] = key[1:] != key[:-1]
    key, vals = key[last], vals[last]
    r = (key // n).astype(np.int32)
    rowptr = np.zeros(m + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(np.bincount(r, minlength=m))
    colidx = (key % n).astype(np.int32)
    if values is not None:
        vals = set_values(colidx.size, values)
    return m, n, rowptr.astype(np.int32), colidx, vals.astype(np.float64)


def write_mtx(path: str, m, n, rowptr, colidx, val):
    rows = np.repeat(np.arange(m), np.diff(rowptr))
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write(f"{m} {n} {len(colidx)}\n")
        for r, c, v in zip(rows, colidx, val):
            f.write(f"{r + 1} {c + 1} {v:.17g}\n")


def transpose_csr(m, n, rowptr, colidx, val):
    """Host CSR -> CSC (= CSR of the transpose), stable. Mirror of reference src/utils.h:161 used by
    tests and by data preparation; the product's device version is tsg_transpose()."""
    order = np.argsort(colidx, kind="stable")
    rows = np.repeat(np.arange(m, dtype=np.int32), np.diff(rowptr))
    colptr = np.zeros(n + 1, dtype=np.int64)
    colptr[1:] = np.cumsum(np.bincount(colidx, minlength=n))
    return colptr.astype(np.int32), rows[order], np.asarray(val)[order]
