"""Wall time of the overlapped end-to-end call (tsg_spgemm_to_host) by slab count, config 2. Run on a B200."""
import sys, os, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
from spgemm_b200 import api, matrices as M
api.init(0)
m, n, rp, ci, v = M.stencil27(int(sys.argv[1]) if len(sys.argv) > 1 else 128)
pin = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in (rp, ci, v)]
def T():
    api.sync(); return time.perf_counter()
a = api.DeviceCSR.upload_ptrs(m, n, pin[0].data_ptr(), pin[1].data_ptr(), pin[2].data_ptr())
ta, tb = api.csr2tile(a, False), api.csr2tile(a, True)
tc, st = api.spgemm(ta, tb); nnz = st["nnzC"]; tc.free()
out = [torch.empty(m + 1, dtype=torch.int32).pin_memory(), torch.empty(nnz, dtype=torch.int32).pin_memory(),
       torch.empty(nnz, dtype=torch.float64).pin_memory()]
gb = nnz * 12 / 1e9
for slabs in (1, 2, 4, 8, 8, 16, 32, 0):
    t0 = T()
    got, st = api.spgemm_to_host(ta, tb, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), nnz, nslabs=slabs)
    t1 = T()
    print("slabs %2d: to_host %.1f ms (%.1f GB/s of C) | steps 1-3 inside %.1f ms" % (slabs, (t1 - t0) * 1e3, gb / (t1 - t0), st["ms_total"]), flush=True)
