"""Times the NT GEMM (STORE epilogue) and the FISTA PROX launch on big shapes; tuning aid."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from decomp_b200 import ops
from decomp_b200._device import empty2d
dev = torch.device('cuda', 0)
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for (M, N, K) in [(131072, 256, 4096), (100000, 256, 256), (400000, 256, 256), (100000, 256, 1024)]:
    A = torch.randn((M, K), dtype=torch.float64, device=dev)
    B = torch.randn((N, K), dtype=torch.float64, device=dev)
    out = empty2d(M, N)
    ms = timeit(lambda: ops.gemm_nt(A, B, ops.epilogue(ops.EPI_STORE, out)))
    print('NT store M=%d N=%d K=%d: %.3f ms %.2f TF/s' % (M, N, K, ms, 2.0 * M * N * K / ms / 1e9))
    del A, B, out
for (K, M, N) in [(131072, 256, 4096), (1000000, 256, 256), (8192, 2048, 1024), (8192, 1024, 4096), (8192, 1024, 1024)]:
    A = torch.randn((K, M), dtype=torch.float64, device=dev)
    B = torch.randn((K, N), dtype=torch.float64, device=dev)
    out = empty2d(M, N)
    ws = ops.gemm_tn_workspace(M, N, K, dev)
    ms = timeit(lambda: ops.gemm_tn(A, B, out, workspace=ws))
    print('TN K=%d M=%d N=%d: %.3f ms %.2f TF/s' % (K, M, N, ms, 2.0 * M * N * K / ms / 1e9))
    del A, B, out, ws
