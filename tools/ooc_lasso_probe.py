"""lasso.solve on a host batch cut into many chunks (1e6 problems, 8.2 GB of y): run time and peak device memory."""
import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch
from decomp_b200 import lasso
rng = np.random.RandomState(0)
B, k, f = 1000000, 256, 1024
A = rng.randn(k, f)
y = np.empty((B, f))
blk = rng.randn(50000, f)
for r in range(0, B, 50000):
    y[r:r + 50000] = blk * (1 + 0.001 * (r // 50000))
torch.cuda.reset_peak_memory_stats()
t0 = time.perf_counter()
it, x = lasso.solve(y, A, 0.1, tol=0.0, method='fista', maxiter=21)
torch.cuda.synchronize()
print('rows %d (y %.1f GB): %.2f s, peak device memory %.1f GB, chunks %d' % (
    B, y.nbytes / 1e9, time.perf_counter() - t0, torch.cuda.max_memory_allocated() / 1e9,
    len(lasso._row_chunks(B, f, k, torch.device('cuda', 0)))))
it2, x2 = lasso.solve(y[:50000], A, 0.1, tol=0.0, method='fista', maxiter=21)
print('first block equals a separate solve:', np.array_equal(x[:50000], x2), 'scaled block relates:',
      float(np.abs(x[50000:100000]).max()), float(np.abs(x2).max()))
