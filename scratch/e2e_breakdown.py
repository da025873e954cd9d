import sys, os, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
from spgemm_b200 import api, matrices as M
api.init(0)
m,n,rp,ci,v = M.stencil27(128)
pin = [torch.from_numpy(np.ascontiguousarray(x)).pin_memory() for x in (rp,ci,v)]
def T():
    api.sync(); return time.perf_counter()
out=None
for it in range(3):
    t0=T(); a = api.DeviceCSR.upload_ptrs(m,n,pin[0].data_ptr(),pin[1].data_ptr(),pin[2].data_ptr())
    t1=T(); ta = api.csr2tile(a, False)
    t2=T(); tb = api.csr2tile(a, True)
    t3=T(); tc, st = api.spgemm(ta, tb)
    t4=T(); cc = api.tile2csr_device(tc)
    t5=T()
    if out is None:
        out=[torch.empty(m+1,dtype=torch.int32).pin_memory(), torch.empty(st['nnzC'],dtype=torch.int32).pin_memory(), torch.empty(st['nnzC'],dtype=torch.float64).pin_memory()]
        t5=T()
    cc.download_into(out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr())
    t6=T()
    for o in (cc,tc,ta,tb,a): o.free()
    t7=T()
    h2d=sum(t.numel()*t.element_size() for t in pin)/1e9; d2h=sum(t.numel()*t.element_size() for t in out)/1e9
    print(it, 'upload %.1f ms (%.1f GB/s) | csr2tile A %.1f | csr2tile B %.1f | spgemm %.1f | tile2csr %.1f | download %.1f ms (%.1f GB/s) | free %.1f | total %.1f' % ((t1-t0)*1e3, h2d/(t1-t0), (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, (t5-t4)*1e3, (t6-t5)*1e3, d2h/(t6-t5), (t7-t6)*1e3, (t7-t0)*1e3))
