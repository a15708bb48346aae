"""Small batched Lasso (reference-test sized): 1000 problems, A (20, 100), 500 FISTA iterations, host arrays in/out."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from decomp_b200 import lasso
from oracle import decomp_oracle as orc
rng = np.random.RandomState(0)
B, k, f = 1000, 20, 100
A = rng.randn(k, f)
y = (rng.randn(B, k) * np.rint(rng.uniform(size=(B, k)))).dot(A) + 0.1 * rng.randn(B, f)
for resident in (True, False):
    lasso.USE_RESIDENT = resident
    for tol in (0.0, 1e-4):
        lasso.solve(y, A, 0.1, tol=tol, method='fista', maxiter=500)
        best = 1e9
        for _ in range(5):
            t0 = time.perf_counter(); it, x = lasso.solve(y, A, 0.1, tol=tol, method='fista', maxiter=500); best = min(best, time.perf_counter() - t0)
        print('resident', resident, 'tol', tol, 'it', it, '%.2f ms' % (best * 1e3))
t0 = time.perf_counter(); it0, x0 = orc.lasso(y, A, 0.1, tol=0.0, method='fista', maxiter=500); print('numpy oracle %.1f ms' % ((time.perf_counter() - t0) * 1e3), float(np.max(np.abs(x - x0))))
