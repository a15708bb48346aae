"""Short TF32-split FISTA run for ncu (BASELINE configs[1] shape)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from decomp_b200 import lasso
dev = torch.device('cuda', 0)
y, A = bench.fista_data_device(torch, 100000, 256, 1024, 0, dev)
solver = lasso.LassoSolver(y, A, 0.1, None, 0.0, 8, 'fista', False, precision='tf32x3')
solver.iterate(0, 8)
st = solver.finish()
torch.cuda.synchronize()
print('ok', float(st.result.abs().sum().item()))
