import glob
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle (test infrastructure) and, if it is missing, the CUDA library (nvcc cross-compiles)."""
    from oracle import oracle as orc
    orc.build()
    so = os.path.join(ROOT, "spgemm_b200", "libtilespgemm_b200.so")
    if not os.path.exists(so):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "spgemm_b200", "csrc")])


def golden_cases():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "ref_*.npz"))
                  if "bitmask_h" not in p)


def general_tile_golden_cases():
    """gtile_<case>_<tile_size_m>x<tile_size_n>.npz: the reference's csr2tile / tile2csr at runtime tile sizes."""
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "gtile_*.npz")))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


TILE_FIELDS = ("tile_ptr", "tile_columnidx", "tile_rowidx", "tile_nnz", "val", "col", "ptr", "mask")


def assert_tiled_equal(got, exp, what, fields=TILE_FIELDS + ("csc_tile_ptr", "csc_tile_rowidx"), val_rtol=0.0):
    """got/exp: dicts or objects with the Tiled field names. Integer fields bit-exact; val exact unless val_rtol."""
    def g(o, k):
        return o[k] if isinstance(o, dict) else getattr(o, k)
    for k in ("m", "n", "tilem", "tilen", "numtile", "nnz"):
        assert int(g(got, k)) == int(g(exp, k)), f"{what}: {k} {g(got, k)} != {g(exp, k)}"
    for k in fields:
        try:
            e = g(exp, k)
        except (KeyError, AttributeError):
            continue
        if e is None:
            continue
        a = np.asarray(g(got, k))
        e = np.asarray(e)
        assert a.shape == e.shape, f"{what}: {k} shape {a.shape} != {e.shape}"
        if k == "val" and val_rtol > 0:
            denom = np.maximum(np.abs(e), 1e-300)
            err = np.max(np.abs(a - e) / denom) if a.size else 0.0
            assert err <= val_rtol, f"{what}: val max rel err {err:g} > {val_rtol:g}"
        else:
            assert np.array_equal(a.astype(np.int64) if a.dtype.kind in "iu" else a,
                                  e.astype(np.int64) if e.dtype.kind in "iu" else e), f"{what}: {k} differs"
