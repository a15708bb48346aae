"""Where does an end-to-end lasso.solve(host arrays) call spend its time? (C2, 200 FISTA iterations)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from decomp_b200 import lasso, ops
from decomp_b200._device import to_device2d, to_host
dev = torch.device('cuda', 0)
B, k, f, K = 100000, 256, 1024, 200
y, A = bench.fista_data_device(torch, B, k, f, 0, dev)
yh = torch.empty((B, f), dtype=torch.float64, pin_memory=True); yh.copy_(y)
Ah = torch.empty((k, f), dtype=torch.float64, pin_memory=True); Ah.copy_(A)
y_np, A_np = yh.numpy(), Ah.numpy()
lasso.solve(y_np, A_np, 0.1, tol=0.0, method='fista', maxiter=K)
def T(): torch.cuda.synchronize(); return time.perf_counter()
for rep in range(2):
    t0 = T(); y2 = to_device2d(y_np, dev, copy=False); A2 = to_device2d(A_np, dev, copy=False); t1 = T()
    s = lasso.LassoSolver(y2, A2, 0.1, None, 0.0, K, 'fista', False); t2 = T()
    tq = time.perf_counter(); s.iterate(0, K); tq = time.perf_counter() - tq; t3 = T()
    st = s.finish(); t4 = T()
    res = to_host(st.result, y_np, np.float64); t5 = T()
    print('H2D %.1f ms | setup %.1f | iterate %.1f (host enqueue %.1f) | finish %.1f | D2H %.1f | total %.1f' % (
        (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, tq*1e3, (t4-t3)*1e3, (t5-t4)*1e3, (t5-t0)*1e3))
t0 = T(); it, x = lasso.solve(y_np, A_np, 0.1, tol=0.0, method='fista', maxiter=K); t1 = T()
print('lasso.solve total %.1f ms' % ((t1 - t0) * 1e3))
# raw copy speeds
d = torch.empty((B, f), dtype=torch.float64, device=dev)
t0 = T(); d.copy_(yh, non_blocking=True); t1 = T(); print('pinned H2D 819 MB: %.1f ms = %.1f GB/s' % ((t1-t0)*1e3, 0.8192/(t1-t0)))
xo = torch.empty((B, k), dtype=torch.float64, device=dev)
t0 = T(); xc = xo.cpu(); t1 = T(); print('pageable D2H 205 MB: %.1f ms = %.1f GB/s' % ((t1-t0)*1e3, 0.2048/(t1-t0)))
xp = torch.empty((B, k), dtype=torch.float64, pin_memory=True)
t0 = T(); xp.copy_(xo, non_blocking=True); t1 = T(); print('pinned D2H 205 MB: %.1f ms = %.1f GB/s' % ((t1-t0)*1e3, 0.2048/(t1-t0)))
