"""Step timings of one workload under whatever TSG_* switches the environment carries.
usage: [TSG_STEP3=rows] python scratch/quick_steps.py stencil27-128 [reps]   (run on a B200)"""
import os
import sys

sys.path.insert(0, os.getcwd())
import bench  # noqa: E402  (workload table only)
from spgemm_b200 import api  # noqa: E402

name = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
gen, aat, _ = bench.WORKLOADS[name]
api.init(0)
m, n, rp, ci, v = gen()
d = api.DeviceCSR.upload(m, n, rp, ci, v)
dB = api.transpose(d) if aat else d
tA, tB = api.csr2tile(d, False), api.csr2tile(dB, True)
best = None
for _ in range(reps):
    if name in bench.SLAB_PAIRS:
        st, _per = api.spgemm_slabs(tA, tB, max_pairs=bench.SLAB_PAIRS[name])
    else:
        c, st = api.spgemm(tA, tB)
        c.free()
    if best is None or st["ms_total"] < best["ms_total"]:
        best = st
print(name, "TSG_STEP3=%s" % os.environ.get("TSG_STEP3", "-"),
      {k: round(best[k], 3) for k in ("ms_step1", "ms_step2", "ms_step3", "ms_alloc", "ms_total")}, "nnzC", best["nnzC"], flush=True)
