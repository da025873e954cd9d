"""TEST INFRASTRUCTURE: runs the emulated general-tile path (tests/emu/gentile_emu.cpp -> spgemm_b200/csrc/gentile.cu as
plain C++) under AddressSanitizer, in a process started with LD_PRELOAD=libasan.so by tests/test_gentile_emu.py: every
load and store of every kernel and of the host orchestration is bounds-checked (the GPU-side tool, compute-sanitizer, is
not available on the GPU pool). Usage: python run_asan.py <libgentile_emu_asan.so>"""
import ctypes as C
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import test_gentile_emu as T  # noqa: E402
from spgemm_b200 import matrices as M  # noqa: E402

lib = C.CDLL(sys.argv[1])
lib.emu_last_error_string.restype = C.c_char_p
lib.emu_free.restype = None
lib.emu_clear_error.restype = None
n = 0
for name in sorted(T.CASES):
    for tile in T.TILE_SIZES:
        m, nn, rp, ci, _ = T.CASES[name]()
        T.run_case(lib, tile[0], tile[1], (m, nn, rp, ci, M.set_values(len(ci), "mod10")))
        n += 1
A, B = M.random_sparse(70, 100, 0.05, seed=21), M.random_sparse(100, 45, 0.06, seed=22)
for tile in ((32, 32), (64, 16), (16, 48)):
    T.run_case(lib, tile[0], tile[1], A, B)
    n += 1
print("asan cases ok:", n)
