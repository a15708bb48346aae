"""End-to-end lasso.solve at C2 from ordinary (pageable) numpy arrays vs page-locked ones."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from decomp_b200 import lasso
dev = torch.device('cuda', 0)
B, k, f, K = 100000, 256, 1024, 200
y, A = bench.fista_data_device(torch, B, k, f, 0, dev)
y_page, A_page = y.cpu().numpy(), A.cpu().numpy()
yh = torch.empty((B, f), dtype=torch.float64, pin_memory=True); yh.copy_(y)
y_pin = yh.numpy()
del y
def T(): torch.cuda.synchronize(); return time.perf_counter()
for name, arr in (('pageable', y_page), ('pinned', y_pin)):
    for mode in ('pipelined', 'one piece'):
        lasso.PIPELINE_MIN_BYTES = (64 << 20) if mode == 'pipelined' else (1 << 60)
        best, x = 1e9, None
        for _ in range(3):
            x = None
            t0 = T(); it, x = lasso.solve(arr, A_page, 0.1, tol=0.0, method='fista', maxiter=K); best = min(best, T() - t0)
        print('%-9s %-10s %.1f ms' % (name, mode, best * 1e3))

from decomp_b200 import _device
d = None
for name, thr in (('staged (4 threads)', 64 << 20), ('torch pageable copy', 1 << 60)):
    _device.STAGE_MIN_BYTES = thr
    best = 1e9
    for _ in range(3):
        d = None
        t0 = T(); d = _device.to_device2d(y_page, dev); best = min(best, T() - t0)
    print('upload 819 MB, %-20s %.1f ms = %.1f GB/s' % (name, best * 1e3, 0.8192 / best))
_device.STAGE_MIN_BYTES = 64 << 20
lasso.PIPELINE_MIN_BYTES = 1 << 60
for tol in (1e-12,):
    best = 1e9
    for _ in range(3):
        t0 = T(); it, x = lasso.solve(y_page, A_page, 0.1, tol=tol, method='fista', maxiter=K); best = min(best, T() - t0)
    print('pageable, tol > 0 (one piece, staged upload): %.1f ms' % (best * 1e3))
