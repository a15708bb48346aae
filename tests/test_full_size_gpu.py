"""Parity at BASELINE.json's full sizes through size-independent properties.

* Batched Lasso (configs[1]: 100 000 problems, A (256, 1024), alpha = 0.1, float64): the problems are independent
  given A, so the rows the GPU returns for a random subset must equal the oracle run on just those rows.
* NMF-MU (configs[2] shape per row block: 4096 features, k = 256): two full sweeps at 131 072 rows against the
  oracle (the CPU needs a few seconds per sweep there; the 1 000 000-row run of the bench differs only in the row
  count, which the ragged-tile kernel tests cover), plus monotone decrease of the objective.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu
RTOL = 1.0e-10


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.mark.parametrize('method,mask1d', [('fista', False), ('ista', False), ('fista', True), ('fista_pos', False)])
def test_lasso_full_size_rows_are_independent(method, mask1d):
    import bench
    from decomp_b200 import lasso
    from oracle import decomp_oracle as orc
    dev = torch.device('cuda', 0)
    B, k, f = 100000, 256, 1024
    y, A = bench.fista_data_device(torch, B, k, f, 3, dev)
    mask = None
    if mask1d:
        mask = (torch.rand(f, dtype=torch.float64, device=dev) > 0.2).double()
    it, x = lasso.solve(y, A, 0.1, tol=0.0, method=method, maxiter=100, mask=mask)
    assert it == 99 and tuple(x.shape) == (B, k) and bool(torch.isfinite(x).all().item())
    rows = np.random.RandomState(0).choice(B, 48, replace=False)
    rows_t = torch.from_numpy(rows).to(dev)
    it0, x_ref = orc.lasso(y[rows_t].cpu().numpy(), A.cpu().numpy(), 0.1, tol=0.0, method=method, maxiter=100,
                           mask=None if mask is None else mask.cpu().numpy())
    assert it0 == it
    assert rel(x[rows_t].cpu().numpy(), x_ref) <= RTOL
    assert float((x != 0).double().mean().item()) > 0.05            # not the trivial solution


def test_nmf_full_width_two_sweeps_vs_oracle():
    import bench
    from decomp_b200 import nmf
    from oracle import decomp_oracle as orc
    dev = torch.device('cuda', 0)
    n, f, k = 131072, 4096, 256
    y, D0, _ = bench.nmf_data_device(torch, n, f, k, 1, dev)
    it, D, x = nmf.solve(y, D0, tol=0.0, maxiter=3)
    yh, Dh = y.cpu().numpy(), D0.cpu().numpy()
    it0, D_ref, x_ref = orc.nmf_mu(yh, Dh, tol=0.0, maxiter=3)
    assert it == it0 == 3
    assert rel(D.cpu().numpy(), D_ref) <= RTOL
    assert rel(x.cpu().numpy(), x_ref) <= RTOL
    # the multiplicative update never increases the objective
    objs = []
    for sweeps in (1, 2, 3):
        _, Ds, xs = nmf.solve(y, D0, tol=0.0, maxiter=sweeps + 1)
        r = y - xs @ Ds
        objs.append(float(0.5 * (r * r).sum().item()))
    assert objs[0] >= objs[1] >= objs[2]
