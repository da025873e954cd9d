"""tile2csr on config 2's C: CUDA-event time of the device conversion (K back-to-back calls), and the csr2tile pair."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from spgemm_b200 import api, matrices as M
api.init(0)
name = sys.argv[1] if len(sys.argv) > 1 else "stencil27-128"
gen = {"stencil27-128": lambda: M.stencil27(128), "stencil27-64": lambda: M.stencil27(64), "blockfem-2M": lambda: M.blockfem(333334)}[name]
m, n, rp, ci, _ = gen()
v = M.set_values(len(ci), "mod10")
d = api.DeviceCSR.upload(m, n, rp, ci, v)
out = {"workload": name}
for label, fn in (("csr2tile_row_major", lambda: api.csr2tile(d, False)), ("csr2tile_col_major", lambda: api.csr2tile(d, True))):
    fn().free()
    api.timer_start()
    for _ in range(5):
        fn().free()
    out[label + "_ms"] = api.timer_stop() / 5
tA, tB = api.csr2tile(d, False), api.csr2tile(d, True)
tC, st = api.spgemm(tA, tB)
api.tile2csr_device(tC).free()
api.timer_start()
for _ in range(5):
    api.tile2csr_device(tC).free()
ms = api.timer_stop() / 5
nnzC, tiles = st["nnzC"], st["numblkC"]
byts = nnzC * 10 + tiles * (32 + 4 + 4) + (m + 1) * 4 + nnzC * 12   # tiled C read (Val, Col, Ptr, tile col, tile nnz) + CSR written
out.update(tile2csr_ms=ms, nnzC=nnzC, C_tiles=tiles, algorithmic_bytes=byts, gbs=byts / ms / 1e6)
r, c, vv = api.tile2csr_device(tC).download()
import scipy.sparse as sp
S = sp.csr_matrix((v, ci, rp), shape=(m, n))
ref = (S @ S).tocsr(); ref.sort_indices()
out["matches_scipy"] = bool(np.array_equal(r, ref.indptr) and np.array_equal(c, ref.indices) and np.array_equal(vv, ref.data)) if ref.nnz == len(c) else "structure differs (explicit zeros?)"
print(json.dumps(out))
