"""Epilogue cost experiments on the FISTA shape (M=100000, N=K=256)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from decomp_b200 import ops
from decomp_b200._device import empty2d
dev = torch.device('cuda', 0)
def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
M, N, K = 100000, 256, 256
A = torch.randn((M, K), dtype=torch.float64, device=dev)
B = torch.randn((N, K), dtype=torch.float64, device=dev) * 0.01
bufs = [torch.randn((M, N), dtype=torch.float64, device=dev) for _ in range(5)]
out, out2, x, other, prev = bufs
colvec = torch.rand(N, dtype=torch.float64, device=dev)
step = torch.full((1,), 0.1, dtype=torch.float64, device=dev)
E = ops.epilogue
cases = {
 'STORE (0 ld, 1 st)': E(ops.EPI_STORE, out),
 'STORE_MASK (1 ld, 1 st)': E(ops.EPI_STORE_MASK, out, mask=other),
 'MU_NUM (2 ld, 1 st)': E(ops.EPI_MU_NUM, out, x=x, other=other),
 'PROX no out2 (3 ld, 1 st)': E(ops.EPI_PROX, out, x=x, other=other, prev=prev, colvec=colvec, colvec2=colvec, step=step, momentum=0.3),
 'PROX (3 ld, 2 st)': E(ops.EPI_PROX, out, out2=out2, x=x, other=other, prev=prev, colvec=colvec, colvec2=colvec, step=step, momentum=0.3),
 'PROX x=A operand (as FISTA)': E(ops.EPI_PROX, out, out2=out2, x=A, other=other, prev=prev, colvec=colvec, colvec2=colvec, step=step, momentum=0.3),
 'PROX in place (out=prev)': E(ops.EPI_PROX, prev, out2=out2, x=A, other=other, prev=prev, colvec=colvec, colvec2=colvec, step=step, momentum=0.3),
}
thr = ops.vector(N, dev); thr.copy_(colvec * 0.1)
cases['PROXQ (2 TMA tiles, 2 st)'] = E(ops.EPI_PROXQ, prev, out2=out2, other=other, prev=prev, colvec=thr, colvec2=colvec, flags=1, momentum=0.3)
for name, epi in cases.items():
    ms = timeit(lambda: ops.gemm_nt(A, B, epi))
    print('%-32s %.3f ms' % (name, ms))
