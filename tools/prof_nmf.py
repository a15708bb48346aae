"""Short NMF-MU run for ncu: configs[2] shape at a reduced row count, a few sweeps.
usage: python tools/prof_nmf.py [rows] [sweeps] [fp64|tf32x3] [f] [k] [masked: 0|1]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from decomp_b200 import nmf
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
sweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
precision = sys.argv[3] if len(sys.argv) > 3 else 'fp64'
f = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
k = int(sys.argv[5]) if len(sys.argv) > 5 else 256
masked = len(sys.argv) > 6 and sys.argv[6] == '1'
dev = torch.device('cuda', 0)
y, D0, mask = bench.nmf_data_device(torch, rows, f, k, 0, dev, masked=masked)
X = torch.ones((rows, k), dtype=torch.float64, device=dev)
solver = nmf.MuSolver(y, D0, X, 0.0, mask=mask, precision=precision)
for it in range(1, sweeps + 1):
    solver.sweep(it)
torch.cuda.synchronize()
print('ok', float(solver.Dbuf[sweeps % 2].sum().item()))
