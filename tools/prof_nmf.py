"""Short NMF-MU run for ncu: 131072 rows x 4096 features, k=256 (the per-GPU shape of BASELINE configs[2] cut in rows)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from decomp_b200 import nmf
dev = torch.device('cuda', 0)
y, D0, _ = bench.nmf_data_device(torch, 131072, 4096, 256, 0, dev)
X = torch.ones((131072, 256), dtype=torch.float64, device=dev)
s = nmf.MuSolver(y, D0, X, 0.0)
for it in range(1, 4):
    s.sweep(it)
torch.cuda.synchronize()
print('ok', float(X.sum().item()))
