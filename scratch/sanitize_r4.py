"""compute-sanitizer memcheck over the kernels added / changed in this session (no torch, no pytest: keeps the run short):
the general-tile path at several tile shapes with both numeric kernels, the pipelined 16x16 k_step3_dense, the split k_plan_build."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc  # noqa: E402
from spgemm_b200 import api, matrices as M  # noqa: E402

api.init(0)


def general(name, gen, tm, tn, mode):
    os.environ["TSG_GT_NUMERIC"] = mode
    m, n, rp, ci, _ = gen
    v = M.set_values(len(ci), "mod10")
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    tA, tB = api.gtile_csr2tile(d, False, tm, tn), api.gtile_csr2tile(d, True, tn, tm)
    tC, st = api.gtile_spgemm(tA, tB)
    csr = api.gtile_tile2csr(tC)
    r, c, vv = csr.download()
    er, ec, ev = orc.spgemm_spa((rp, ci, v), (rp, ci, v), n)
    assert np.array_equal(r, er) and np.array_equal(c, ec) and np.array_equal(vv, ev), name
    print("ok", name, tm, tn, mode, st["nnzC"], st["tiles_dense"], flush=True)
    for o in (csr, tC, tA, tB, d):
        o.free()


def tuned(name, gen, env):
    for k, val in env.items():
        os.environ[k] = val
    m, n, rp, ci, _ = gen
    v = M.set_values(len(ci), "mod10")
    r, c, vv, st = api.spgemm_csr_host(m, n, n, (rp, ci, v))
    er, ec, ev = orc.spgemm_spa((rp, ci, v), (rp, ci, v), n)
    assert np.array_equal(r, er) and np.array_equal(c, ec) and np.array_equal(vv, ev), name
    print("ok", name, env, st["nnzC"], st["tiles_dense"], st["plan_recipes"], flush=True)
    for k in env:
        del os.environ[k]


for tm, tn in ((32, 32), (32, 64), (32, 16)):
    for mode in ("dense", "gather"):
        general("blockfem", M.blockfem(50), tm, tn, mode)
        general("stencil", M.stencil27(8, 7, 5), tm, tn, mode)
general("rmat", M.rmat(9, 6, seed=3), 48, 64, "gather")
general("ragged", M.random_sparse(203, 203, 0.03, seed=11), 128, 16, "gather")
general("empty", (33, 33, np.zeros(34, np.int32), np.zeros(0, np.int32), np.zeros(0)), 32, 32, "dense")
tuned("blockfem16 dense", M.blockfem(60), {"TSG_STEP3": "dense", "TSG_PLANS": "0"})
tuned("full48 dense", M.random_sparse(48, 48, 5.0, seed=15), {"TSG_STEP3": "dense", "TSG_PLANS": "0"})
tuned("stencil plans", M.stencil27(32, 4, 4), {"TSG_PLANS": "2"})
tuned("lap2d plans", M.lap2d(40), {"TSG_PLANS": "2"})
print("sanitize_r4: all cases ok")
