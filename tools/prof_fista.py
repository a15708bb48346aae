"""Short FISTA run for ncu: full-size batch (BASELINE configs[1]), a handful of iterations."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from decomp_b200 import lasso
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 12
dev = torch.device('cuda', 0)
y, A = bench.fista_data_device(torch, 100000, 256, 1024, 0, dev)
solver = lasso.LassoSolver(y, A, 0.1, None, 0.0, iters, 'fista', False)
solver.iterate(0, iters)
st = solver.finish()
torch.cuda.synchronize()
print('ok', float(st.result.abs().sum().item()))
