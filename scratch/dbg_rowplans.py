import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TSG_DEBUG"] = "1"
import numpy as np
from spgemm_b200 import api, matrices as M
api.init(0)
def run(name, gen, reps=("1", "0")):
    m, n, rp, ci, _ = gen
    v = M.set_values(len(ci), "mod10")
    d = api.DeviceCSR.upload(m, n, rp, ci, v)
    tA, tB = api.csr2tile(d, False), api.csr2tile(d, True)
    for rpn in reps:
        os.environ["TSG_ROWPLANS"] = rpn
        tC, st = api.spgemm(tA, tB)
        print(name, "rowplans", rpn, {k: st[k] for k in ("numblkC", "nnzC", "pairs", "plan_recipes", "row_templates")}, flush=True)
        tC.free()
    for o in (tA, tB, d):
        o.free()
small = len(sys.argv) > 1
run("rmat s12 (garbage for the arenas)", M.rmat(12, 8, seed=3), reps=("1",))
run("lap2d_96", M.lap2d(96))
run("stencil27_17 (no templates)", M.stencil27(17), reps=("1",))
run("stencil27_32x16x16", M.stencil27(32, 16, 16))
if not small:
    run("rmat s14", M.rmat(14, 8, seed=3), reps=("1",))
    run("stencil27_64", M.stencil27(64))
