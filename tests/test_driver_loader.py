"""The C driver's MatrixMarket loader (driver/main.c load_mtx; the reference's is src/mmio_highlevel.h:593-759) without a
device: field types, symmetry, sort + dedupe vs the reference's file order, the binary cache."""
import os
import re
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT

EXE = os.path.join(ROOT, "driver", "test_b200")


def run(path, **env):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "driver")])
    e = dict(os.environ, TSG_DRIVER_PARSE_ONLY="1", **{k: str(v) for k, v in env.items()})
    out = subprocess.run([EXE, "-d", "0", "-aat", "1", str(path), "16", "16"], capture_output=True, text=True, env=e, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    m = re.search(r"parsed: m=(\d+) n=(\d+) nnz=(\d+) symmetric=(\d) sorted=(\d) index_checksum=(-?\d+) value_sum=([-\d.]+)", out.stdout)
    assert m, out.stdout
    return dict(m=int(m[1]), n=int(m[2]), nnz=int(m[3]), sym=int(m[4]), sorted=int(m[5]), chk=int(m[6]), vsum=float(m[7]), out=out.stdout)


def checksum(S):
    S = S.tocoo()
    return int(((S.row.astype(np.int64) + 1) * (S.col.astype(np.int64) + 1)).sum())


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "spgemm_b200", "libtilespgemm_b200.so")), reason="library not built")
def test_mtx_loader_fields_symmetry_cache(tmp_path):
    rng = np.random.default_rng(3)
    S = sp.random(40, 37, density=0.08, random_state=5, format="coo")
    entries = list(zip(S.row + 1, S.col + 1, rng.integers(1, 9, S.nnz)))
    rng.shuffle(entries)                                      # file order is not row order
    entries += entries[:7]                                    # and holds duplicates
    for field, fmt in (("real", "{} {} {}.5"), ("integer", "{} {} {}"), ("complex", "{} {} {}.0 2.5"), ("pattern", "{} {}")):
        p = tmp_path / f"g_{field}.mtx"
        with open(p, "w") as f:
            f.write(f"%%MatrixMarket matrix coordinate {field} general\n% comment\n40 37 {len(entries)}\n")
            for r, c, v in entries:
                f.write(fmt.format(r, c, v) + "\n")
        a = run(p, TSG_MTX_CACHE=0)
        assert (a["m"], a["n"], a["nnz"], a["sym"], a["sorted"]) == (40, 37, S.nnz, 0, 1), a     # sorted, duplicates merged
        assert a["chk"] == checksum(S)
        raw = run(p, TSG_MTX_CACHE=0, TSG_MTX_RAW=1)                                             # the reference's file order
        assert raw["nnz"] == len(entries) and raw["sorted"] == 0
    # symmetric: mirrored off-diagonal entries
    L = sp.tril(sp.random(30, 30, density=0.1, random_state=9, format="csr") + sp.identity(30)).tocoo()
    p = tmp_path / "sym.mtx"
    with open(p, "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate real symmetric\n30 30 {L.nnz}\n")
        for r, c, v in zip(L.row + 1, L.col + 1, L.data):
            f.write(f"{r} {c} {v}\n")
    a = run(p, TSG_MTX_CACHE=0)
    full = (L + sp.tril(L, -1).T).tocsr()
    assert (a["nnz"], a["sym"]) == (full.nnz, 1) and a["chk"] == checksum(full)
    # binary cache: written on the first load, used on the second, same matrix
    first = run(p)
    assert "binary cache" not in first["out"] and os.path.exists(str(p) + ".tsgcsr")
    second = run(p)
    assert "binary cache" in second["out"] and (second["nnz"], second["chk"], second["vsum"]) == (first["nnz"], first["chk"], first["vsum"])
