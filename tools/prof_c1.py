"""BASELINE configs[0] through the one-launch small-problem kernel, for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import golden_cases as gc
from decomp_b200 import nmf
y, D0, mask = gc._nmf_data(1000, 200, 20, 0, 'l2', reference_order=False)
for m in (None, mask):
    it, D, x = nmf.solve(y, D0.copy(), tol=0.0, maxiter=101, mask=m)
torch.cuda.synchronize()
print('ok', float(D.sum()))
