"""Phase timing of one dictionary-learning minibatch step at C4 scale (complex128, f=2048, k=512, minibatch 8192)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from decomp_b200 import ops
from decomp_b200._device import empty2d, zeros2d
from decomp_b200._lib import rview
from decomp_b200.lasso import lasso_device
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev); g.manual_seed(0)
def crandn(*s): return torch.complex(torch.randn(s, dtype=torch.float64, device=dev, generator=g), torch.randn(s, dtype=torch.float64, device=dev, generator=g))
mb, f, k = 8192, 2048, 512
D = crandn(k, f); D = D / D.abs().pow(2).sum(-1, keepdim=True).sqrt()
x = crandn(mb, k) * torch.rand((mb, k), dtype=torch.float64, device=dev, generator=g)
y = x @ D + 0.1 * crandn(mb, f)
mask = (torch.rand((mb, f), dtype=torch.float64, device=dev, generator=g) > 0.1).double()
def timed(name, fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    print('%-44s %8.3f ms' % (name, e0.elapsed_time(e1) / reps))
xm = x.clone()
timed('lasso fista x10 unmasked', lambda: lasso_device(y, D, 0.1, xm, 1e-5, 10, 'fista', False, None, out=xm))
timed('lasso fista x10 masked', lambda: lasso_device(y, D, 0.1, xm, 1e-5, 10, 'fista', False, mask, out=xm))
S, T = zeros2d(k, k, True, dev), zeros2d(k, f, True, dev)
ws = ops.gemm_tn_workspace_for([(2 * k, 2 * k, mb), (2 * k, 2 * f, mb), (f, 2 * k, mb)], dev)
xr = rview(x)
timed('stats S = x^H x', lambda: ops.gemm_tn(xr, xr, rview(S), combine=3, beta=0.5, workspace=ws))
timed('stats T = x^H y', lambda: ops.gemm_tn(xr, rview(y), rview(T), combine=3, beta=0.5, workspace=ws))
Dn = D.clone()
timed('dl_sweep (512 atoms, Gauss-Seidel)', lambda: ops.dl_sweep(rview(S), rview(T), rview(Dn), True))
W = empty2d(mb, k, True, dev)
Sb = torch.zeros((k, f, 2 * k), dtype=torch.float64, device=dev)
def masked_stats(n_atoms=16):
    for a in range(n_atoms):
        ops.dl_atom_weighted(xr, True, a, rview(W))
        ops.gemm_tn(mask, rview(W), Sb[a], combine=1, beta=0.5, workspace=ws)
timed('masked stats, 16 of 512 atoms', masked_stats)
Dt_ws = torch.empty(f * k * 2, dtype=torch.float64, device=dev)
Do = empty2d(k, f, True, dev)
timed('dl_masked_update (streams 8.6 GB)', lambda: ops.dl_masked_update(Sb, rview(T), rview(D), rview(Do), True, Dt_ws))
# ---- pieces of the packed masked statistics
from decomp_b200.dictionary_learning import _pair_chunks
chunks = _pair_chunks(k, 512, dev)
colA, colB = chunks[3]
Xt = empty2d(2 * k, mb, False, dev); Mt = empty2d(f, mb, False, dev)
Wt = empty2d(1024, mb, False, dev); Pt = empty2d(f, 512, True, dev)
timed('transpose x', lambda: ops.make_rhs(xr, False, False, out=Xt))
timed('transpose mask', lambda: ops.make_rhs(mask, False, False, out=Mt))
timed('pair_products_t (512 pairs)', lambda: ops.dl_pair_products_t(Xt, True, colA, colB, Wt))
timed('NT gemm f x 1024, K=8192', lambda: ops.gemm_nt(Mt, Wt, ops.epilogue(ops.EPI_STORE, rview(Pt))))
timed('scatter_stats', lambda: ops.dl_scatter_stats(rview(Pt), True, colA, colB, k, 0.5, Sb))
timed('mirror', lambda: ops.dl_mirror(Sb, k, f, True))
