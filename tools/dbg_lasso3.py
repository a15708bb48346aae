import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import golden_cases as gc
from decomp_b200 import ops, lasso
from oracle import decomp_oracle as orc
SYNC = os.environ.get('DBG_SYNC') == '1'
if SYNC:
    orig = ops.gemm_nt
    def g(*a, **k):
        orig(*a, **k); torch.cuda.synchronize()
    ops.gemm_nt = g
def rel(a, b): return np.max(np.abs(a-b))/max(np.max(np.abs(b)),1e-300)
case = gc.lasso_cases()['fix_ista_nomask']
for mi in (1, 2, 3, 5, 10, 57):
    it, x = lasso.solve(case['y'], case['A'], case['alpha'], tol=0.0, method='ista', maxiter=mi)
    it0, x0 = orc.lasso(case['y'], case['A'], case['alpha'], tol=0.0, method='ista', maxiter=mi)
    print('sync', SYNC, 'maxiter', mi, it, it0, rel(x, x0))
