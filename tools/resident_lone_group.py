"""How fast does ONE warp group of the resident Lasso kernel run when it has the tensor pipe to itself?
Width K (argv[1], default 256): a row block has 8192 / K rows, half per group.  B = 148 x half: every CTA has one row
block with only group 0's half populated (group 1 just walks the ring); B = 148 x 2 half: both halves.  Pipe-bound and perfectly fed, the second takes twice as long per iteration."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from decomp_b200 import lasso, ops

dev = torch.device('cuda', 0)
sms = torch.cuda.get_device_properties(0).multi_processor_count
res = {}
K = int(sys.argv[1]) if len(sys.argv) > 1 else 256          # problem width (atoms): 8192 / K rows per row block
half = 8192 // K // 2
for rows_per_cta in (half, 2 * half, 4 * half, 8 * half):
    B = sms * rows_per_cta
    y, A = bench.fista_data_device(torch, B, K, 1024, 0, dev)
    s = lasso.LassoSolver(y, A, 0.1, None, 0.0, 100000, 'fista', False)
    s.iterate(0, 32)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for r in range(5):
        e0.record()
        s.iterate(32 + 64 * r, 96 + 64 * r)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 64)
    cycles = best * 1e-3 * 1.965e9
    res[rows_per_cta] = {'us_per_iteration': best * 1e3, 'cycles_per_iteration': cycles,
                         'dmma_issue_cycles': rows_per_cta * K * K * 2 / 512 / 4 * 16,
                         'k_blocks': K // 16}
print(json.dumps(res, indent=1))
