"""Small end-to-end case for compute-sanitizer memcheck: every kernel family runs (uniform/fused path, pair path with
heavy rows, dense numeric, gather numeric, col-major conversion, transposition, tile2csr, rowsums, slabs)."""
import os, sys
import numpy as np
sys.path.insert(0, os.getcwd())
from spgemm_b200 import api, matrices as M
api.init(0)
for name, (m, n, rp, ci, v), aat in (("stencil", M.stencil27(9), False), ("fem", M.blockfem(60), False),
                                       ("rmat", M.rmat(11, 16, seed=3), True), ("ragged", M.random_sparse(203, 203, 0.03, seed=1), False)):
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    dB = api.transpose(d) if aat else d
    tA, tB = api.csr2tile(d, False), api.csr2tile(dB, True)
    tot, per = api.spgemm_slabs(tA, tB, max_pairs=20000, sink=lambda c, st: (api.tile_rowsums(c), api.tile2csr_device(c).free()))
    for force in ("dense", "gather", "dmma"):
        os.environ["TSG_STEP3"] = force
    tC, st = api.spgemm(tA, tB)
    tC.download(); tC.free()
    print(name, "ok", tot["nnzC"], st["nnzC"], len(per))
    for o in (tA, tB, d):
        o.free()
    if aat:
        dB.free()
