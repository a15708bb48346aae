"""Device-time measurements of the other BASELINE.json configs (parity cases, not bench lines): C4 dictionary
learning (complex128, 10 % mask) and C5 masked NMF / masked FISTA; sizes scaled by --scale to bound the run."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from decomp_b200 import ops, lasso, nmf, dictionary_learning as dl

ap = argparse.ArgumentParser()
ap.add_argument('--which', default='c5nmf,c5lasso,c4dl,c4dl_nomask,c2ista')
ap.add_argument('--rows', type=int, default=262144)
args = ap.parse_args()
dev = torch.device('cuda', 0)
peak = ops.probe_dmma_tflops()
out = {}

def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

g = torch.Generator(device=dev); g.manual_seed(0)
def randn(*s, cplx=False):
    if cplx:
        return torch.complex(torch.randn(s, dtype=torch.float64, device=dev, generator=g),
                             torch.randn(s, dtype=torch.float64, device=dev, generator=g))
    return torch.randn(s, dtype=torch.float64, device=dev, generator=g)

if 'c5nmf' in args.which:
    n, f, k = args.rows, 1024, 128
    Dt = randn(k, f).clamp_(min=0); y = (randn(n, k).clamp_(min=0) @ Dt + 0.1 * randn(n, f))
    D0 = (Dt + 0.3 * randn(k, f)).clamp_(min=0.1)
    mask = (torch.rand((n, f), dtype=torch.float64, device=dev, generator=g) > 0.1).double()
    X = torch.ones((n, k), dtype=torch.float64, device=dev)
    s = nmf.MuSolver(y, D0, X, 0.0, mask=mask)
    it = [0]
    def sweep():
        it[0] += 1; s.sweep(it[0])
    ms = timed(sweep, 5)
    fl = 12.0 * n * k * f
    out['c5_masked_nmf_sweep'] = dict(rows=n, f=f, k=k, ms=ms, tflops=fl / ms / 1e9, frac=fl / ms / 1e9 / peak,
                                      note='12nkf flop per sweep (no re-association under a mask)')
    del s, y, mask, X

if 'c5lasso' in args.which:
    n, f, k = args.rows, 1024, 128
    A = randn(k, f); y = randn(n, k) @ A + 0.1 * randn(n, f)
    mask = (torch.rand((n, f), dtype=torch.float64, device=dev, generator=g) > 0.1).double()
    s = lasso.LassoSolver(y, A, 0.1, None, 0.0, 100000, 'fista', False, mask=mask)
    it = [0]
    def step():
        s.iterate(it[0], it[0] + 1); it[0] += 1
    ms = timed(step, 10)
    fl = 4.0 * n * k * f
    out['c5_masked_fista_iter'] = dict(rows=n, f=f, k=k, ms=ms, tflops=fl / ms / 1e9, frac=fl / ms / 1e9 / peak,
                                       note='4Bkf flop per iteration (two GEMMs)')
    del s, y, mask

for name, masked in (('c4dl', True), ('c4dl_nomask', False)):
    if name not in args.which.split(','):
        continue
    n, f, k, mb = 32768, 2048, 512, 8192
    Dt = randn(k, f, cplx=True)
    xt = randn(n, k, cplx=True) * torch.rand((n, k), dtype=torch.float64, device=dev, generator=g)
    y = xt @ Dt + 0.1 * randn(n, f, cplx=True)
    D0 = Dt + 0.2 * randn(k, f, cplx=True)
    mask = (torch.rand((n, f), dtype=torch.float64, device=dev, generator=g) > 0.1).double() if masked else None
    del xt
    dl.solve(y[:2 * mb], D0, 0.1, tol=0.0, minibatch=mb, maxiter=2, lasso_method='fista', lasso_iter=10,
             mask=mask[:2 * mb] if masked else None, random_seed=0)          # warm the allocator and the kernels
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    itn, D, x = dl.solve(y, D0, 0.1, tol=0.0, minibatch=mb, maxiter=2, lasso_method='fista', lasso_iter=10, mask=mask,
                         random_seed=0)
    e1.record(); torch.cuda.synchronize()
    steps = n // mb
    ms = e0.elapsed_time(e1) / steps
    per_row = (10 * 16 * k * f + 8 * k * f) if masked else (10 * 8 * k * k + 8 * k * f)
    stat = 2.0 * k * k * f * mb if masked else 8.0 * k * k * mb + 8.0 * k * f * mb
    fl = per_row * mb + stat + 8.0 * k * k * f
    out['c4_dl_step_' + ('masked' if masked else 'unmasked')] = dict(
        minibatch=mb, f=f, k=k, ms_per_minibatch_step=ms, tflops=fl / ms / 1e9, frac=fl / ms / 1e9 / peak,
        finite=bool(torch.isfinite(torch.view_as_real(D)).all().item()),
        note='one epoch of %d minibatch steps incl. shuffle; flops = lasso + statistics (Hermitian half: 2 k^2 f per row) '
             '+ atom update' % steps)
    del y, mask, D, x

if 'c2ista' in args.which:
    B, k, f = 100000, 256, 1024
    A = randn(k, f); y = randn(B, k) @ A + 0.1 * randn(B, f)
    for rule in ('ista', 'fista'):
        s = lasso.LassoSolver(y, A, 0.1, None, 1e-12, 100000, rule, False)
        s.iterate(0, 1)
        it = [1]
        def step():
            s.iterate(it[0], it[0] + 50); it[0] += 50
        ms = timed(step, 4) / 50
        out['c2_%s_iter_with_checks' % rule] = dict(ms=ms, note='tol > 0: launches of 10 iterations, the last one of each evaluates the convergence test')
        del s
print(json.dumps(out, indent=1))
