import sys, os, time, numpy as np
sys.path.insert(0, os.getcwd())
from spgemm_b200 import api, matrices as M
api.init(0)
m,n,rp,ci,v = M.rmat(16,16,seed=1)
d = api.DeviceCSR.upload(m,n,rp,ci,v); dT = api.transpose(d)
tA, tB = api.csr2tile(d, False), api.csr2tile(dT, True)
w = api.tilerow_weights(tA, tB)
for it in range(3):
    t0=time.perf_counter()
    tot, per = api.spgemm_slabs(tA, tB, max_pairs=1<<26, weights=w)
    dt=(time.perf_counter()-t0)*1e3
    print(it, 'wall %.1f ms'%dt, {k: round(tot[k],2) for k in ('ms_step1','ms_step2','ms_step3','ms_alloc','ms_total')}, 'slabs', tot['slabs'])
    if it==2:
        for p in per: print('   slab', p['trow0'], p['trow1'], 'pairs', p['pairs'], 'tiles', p['numblkC'], 'nnz', p['nnzC'], {k: round(p[k],2) for k in ('ms_step1','ms_step2','ms_step3','ms_alloc')})
