"""csrc/pair_sort.h on the CPU: the in-place heap sort that puts the long pair lists of hub tile-rows (k_s1_heavy) into
ascending-A-tile order -- the serial SPA's summation order (oracle/spa_ref.c, reference src/spgemm_serialref_spa_new.h) --
so that C's values are reproducible run to run. Exactly the device function, compiled by g++."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

SRC = os.path.join(ROOT, "tests", "emu", "pair_sort_host.cpp")
HDR = os.path.join(ROOT, "spgemm_b200", "csrc", "pair_sort.h")
SO = os.path.join(ROOT, "tests", "emu", "libpair_sort_host.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(SO) or any(os.path.getmtime(SO) < os.path.getmtime(d) for d in (SRC, HDR)):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-Wall", "-Werror", SRC, "-o", SO])
    return C.CDLL(SO)


def _sort(lib, ka, kb):
    pad = 8  # guard words either side: the sort must stay inside [0, n)
    a = np.full(len(ka) + 2 * pad, -7, np.int32)
    b = np.full(len(kb) + 2 * pad, -9, np.int32)
    a[pad:pad + len(ka)] = ka
    b[pad:pad + len(kb)] = kb
    lib.host_pair_heap_sort(C.c_void_p(a.ctypes.data + 4 * pad), C.c_void_p(b.ctypes.data + 4 * pad), C.c_int(len(ka)))
    assert np.all(a[:pad] == -7) and np.all(a[pad + len(ka):] == -7) and np.all(b[:pad] == -9) and np.all(b[pad + len(kb):] == -9)
    return a[pad:pad + len(ka)], b[pad:pad + len(kb)]


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 63, 64, 65, 66, 100, 127, 128, 129, 1000, 4097, 50000])
def test_heap_sort_orders_by_a_tile_and_keeps_the_pairs(lib, n):
    rng = np.random.default_rng(n)
    for trial in range(4 if n < 5000 else 1):
        ka = rng.permutation(3 * n + 5)[:n].astype(np.int32) + 17        # distinct keys, as in a pair list
        kb = rng.integers(0, 1 << 30, n).astype(np.int32)
        if trial == 1:
            ka = np.sort(ka)                                              # already sorted
        if trial == 2:
            ka = np.sort(ka)[::-1].copy()                                 # reversed
        sa, sb = _sort(lib, ka, kb)
        order = np.argsort(ka, kind="stable")
        assert np.array_equal(sa, ka[order]) and np.array_equal(sb, kb[order])


def test_heap_sort_with_equal_keys_is_a_permutation(lib):
    """Keys of one list are distinct in the product path; with equal keys the result is still sorted and still the same
    multiset of pairs (no element lost or duplicated)."""
    rng = np.random.default_rng(5)
    ka = rng.integers(0, 20, 500).astype(np.int32)
    kb = np.arange(500, dtype=np.int32)
    sa, sb = _sort(lib, ka, kb)
    assert np.all(np.diff(sa) >= 0) and np.array_equal(np.sort(sb), kb) and np.array_equal(ka[sb], sa)
