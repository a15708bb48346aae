"""Non-negative matrix factorisation by full-batch multiplicative updates on the B200.

Drop-in for the reference's ``decomp.nmf.solve`` (decomp/nmf.py:16-78 -> nmf_methods/batch_mu.py:8-26
-> nmf_methods/grads.py:77-160): same signature, defaults (``x = ones``), validation and return tuple
``(it, D, x)`` with ``it == maxiter`` on exhaustion.

    y [n, f] ~ x [n, k] . D [k, f],   x, D >= 0,   rows of D have unit L2 norm

One sweep on the device (FP64):

  unmasked 'l2'   G = D D^T ; neg = x G ; x <- x * max(y D^T, 0) / max(neg, eps)      (ratio fused in the GEMM epilogue)
                  T = x^T y ; S = x^T x   (split along the sample axis, deterministic reduction; the only
                  quantities a multi-GPU run all-reduces) ; D <- D * max(T, 0) / max(S D, eps) (fused)
  masked 'l2'     f = (x D) * mask (mask fused) ; neg = f D^T ; pos = (y*mask) D^T ; ... as grads.py:112-125
  'kl'            r = (y*mask) / (x D + eps) (fused) ; pos = r D^T ; neg = row sums of D or mask D^T ; ...
  then            D <- l2_strict(D) and max|D - D_new| < tol  in one kernel that sets a device latch; every
                  later launch checks the latch first, so the host polls it only every POLL_EVERY sweeps.

The re-association x (D D^T) = (x D) D^T is exact in real arithmetic and agrees with the reference to
~1e-15 relative per sweep (rounding order only); it is not available under a mask, where the reference's
own order is kept.
"""
import numpy as np
import torch

from . import ops
from ._device import array_kind, empty2d, full2d, is_torch, np_dtype, require_cuda, to_device2d, to_host
from .utils import assertion

BATCH_METHODS = ['mu']
MINIBATCH_METHODS = ['asg-mu', 'gsg-mu', 'asag-mu', 'gsag-mu', 'svrmu', 'svrmu-acc']
POLL_EVERY = 25


def solve(y, D, x=None, tol=1.0e-3, minibatch=None, maxiter=1000, method='mu', likelihood='l2', mask=None,
          random_seed=None, group=None, **kwargs):
    """NMF, see the module docstring. ``likelihood``: 'l2' | 'gaussian' | 'kl' | 'poisson'.

    ``group``: optional ``torch.distributed`` process group. Each rank passes its own contiguous block of
    rows of ``y`` / ``x`` / ``mask`` and the same ``D``; per sweep only the [k, f] and [k, k] statistics
    are all-reduced. Every rank returns the same ``it`` and ``D`` and its own rows of ``x``.
    """
    array_kind(y, D, x, mask)
    if x is None:
        if is_torch(y):
            x = torch.ones((y.shape[0], D.shape[0]), dtype=y.dtype, device=y.device)
        else:
            x = np.ones((y.shape[0], D.shape[0]), dtype=y.dtype)

    assertion.assert_dtypes(y=y, D=D, x=x)
    assertion.assert_dtypes(y=y, D=D, x=x, mask=mask, dtypes='f')
    assertion.assert_shapes('x', x, 'D', D, axes=1)
    assertion.assert_shapes('y', y, 'D', D, axes=[-1])
    assertion.assert_shapes('y', y, 'mask', mask)
    assertion.assert_ndim('y', y, 2)
    assertion.assert_ndim('D', D, 2)
    assertion.assert_ndim('x', x, 2)
    assertion.assert_nonnegative(D)
    assertion.assert_nonnegative(x)
    if likelihood in ['kl']:
        assertion.assert_nonnegative(y)

    if minibatch is not None:
        raise NotImplementedError('NMF with {} algorithm is not yet implemented.'.format(method)
                                  if method not in MINIBATCH_METHODS else
                                  'minibatch NMF ({}) is outside the B200 hot path; use minibatch=None, '
                                  "method='mu'.".format(method))
    if method != 'mu':
        raise NotImplementedError('Batch-NMF with {} algorithm is not yet implemented.'.format(method))
    if kwargs:
        raise TypeError('solve() got unexpected keyword arguments ' + str(sorted(kwargs)))
    if likelihood in ('l2', 'gaussian'):
        kl = False
    elif likelihood in ('kl', 'poisson'):
        kl = True
    else:
        raise NotImplementedError('Likelihood {} is not implemented for nmf'.format(likelihood))

    device = require_cuda()
    out_dtype = np_dtype(y)
    yd = to_device2d(y, device, copy=False)
    md = to_device2d(mask, device, copy=False) if mask is not None else None
    Dd = to_device2d(D, device, copy=True)
    xd = to_device2d(x, device, copy=True)       # updated in place: always our own copy
    it, Dd, xd = mu_device(yd, Dd, xd, float(tol), int(maxiter), kl, md, group=group)
    return it, to_host(Dd, y, out_dtype), to_host(xd, y, out_dtype)


def mu_device(y, D0, X, tol, maxiter, kl=False, mask=None, group=None):
    """Full-batch MU on device tensors; ``X`` [n, k] is updated in place. Returns ``(it, D, X)``."""
    dev = y.device
    n, f = y.shape
    k = D0.shape[0]
    dist = torch.distributed if group is not None else None

    Dbuf = [empty2d(k, f, False, dev), empty2d(k, f, False, dev)]
    ops.normalize_rows(D0, Dbuf[0], False, True)                     # nmf.py:70
    Draw = empty2d(k, f, False, dev)
    Dt = empty2d(f, k, False, dev)
    POS = empty2d(k, f, False, dev)
    NEG = empty2d(n, k, False, dev)
    ws = ops.gemm_tn_workspace_for([(k, f, n), (k, k, n)], dev)
    checks = tol > 0.0
    latch = torch.zeros(1, dtype=torch.int32, device=dev) if checks else None
    scratch = torch.zeros(1, dtype=torch.int32, device=dev)
    maxdiff = torch.zeros(2, dtype=torch.float64, device=dev)

    ym = y
    if mask is not None:
        ym = empty2d(n, f, False, dev)
        ops.mask_mul(y, mask, ym)                                     # y * mask, once (grads.py:113,123)
    if mask is not None or kl:
        F = empty2d(n, f, False, dev)                                 # the [n, f] intermediate
        NEGD = empty2d(k, f, False, dev)
    else:
        G = empty2d(k, k, False, dev)
        S = empty2d(k, k, False, dev)
    if kl and mask is None:
        ones_kf = full2d(k, f, 1.0, False, dev)
        dsum = torch.empty(k, dtype=torch.float64, device=dev)
        xsum = torch.empty(k, dtype=torch.float64, device=dev)
        dsum_row = dsum.view(1, k)

    def E(kind, out, **kw):
        return ops.epilogue(kind, out, **kw)

    it_done = maxiter
    stopped_at = 0
    for it in range(1, maxiter):
        if checks and it % POLL_EVERY == 0:
            fired = int(latch.item())
            if fired:
                stopped_at = fired
                break
        D, Dn = Dbuf[(it - 1) % 2], Dbuf[it % 2]
        if not kl and mask is None:
            # ---- x update (grads.py:108-111 with f.dot(d.T) re-associated)
            ops.gemm_nt(D, D, E(ops.EPI_STORE, G), skip=latch)
            ops.gemm_nt(X, G, E(ops.EPI_STORE, NEG), skip=latch)
            ops.gemm_nt(y, D, E(ops.EPI_MU_NUM, X, x=X, other=NEG), skip=latch)
            # ---- D update (grads.py:117-121): sufficient statistics over the sample axis
            ops.gemm_tn(X, y, POS, workspace=ws, skip=latch)
            ops.gemm_tn(X, X, S, workspace=ws, skip=latch)
            if dist is not None:
                _allreduce2d(POS, group)
                _allreduce2d(S, group)
            ops.make_rhs(D, False, False, out=Dt, skip=latch)
            ops.gemm_nt(S, Dt, E(ops.EPI_MU_DEN, Draw, x=D, other=POS), skip=latch)
        else:
            ops.make_rhs(D, False, False, out=Dt, skip=latch)
            if not kl:
                # ---- masked l2 (grads.py:112-115, 122-125)
                ops.gemm_nt(X, Dt, E(ops.EPI_STORE_MASK, F, mask=mask), skip=latch)
                ops.gemm_nt(F, D, E(ops.EPI_STORE, NEG), skip=latch)
                ops.gemm_nt(ym, D, E(ops.EPI_MU_NUM, X, x=X, other=NEG), skip=latch)
                ops.gemm_nt(X, Dt, E(ops.EPI_STORE_MASK, F, mask=mask), skip=latch)
                ops.gemm_tn(X, ym, POS, workspace=ws, skip=latch)
                ops.gemm_tn(X, F, NEGD, workspace=ws, skip=latch)
            else:
                # ---- Poisson / KL (grads.py:142-160)
                ops.gemm_nt(X, Dt, E(ops.EPI_KL_RATIO, F, other=y, mask=mask), skip=latch)
                if mask is None:
                    ops.row_sums(D, 1.0, out=dsum)
                    neg_x = E(ops.EPI_MU_NUM, X, x=X, other=dsum_row)
                    neg_x.ldother = 0                                  # one [1, k] row for every sample
                    ops.gemm_nt(F, D, neg_x, skip=latch)
                else:
                    ops.gemm_nt(mask, D, E(ops.EPI_STORE, NEG), skip=latch)
                    ops.gemm_nt(F, D, E(ops.EPI_MU_NUM, X, x=X, other=NEG), skip=latch)
                ops.gemm_nt(X, Dt, E(ops.EPI_KL_RATIO, F, other=y, mask=mask), skip=latch)
                ops.gemm_tn(X, F, POS, workspace=ws, skip=latch)
                if mask is None:
                    ops.col_sums(X, 1.0, out=xsum)
                    if dist is not None:
                        dist.all_reduce(xsum, group=group)
                    ops.scale(ones_kf, NEGD, rowscale=xsum)
                else:
                    ops.gemm_tn(X, mask, NEGD, workspace=ws, skip=latch)
            if dist is not None:
                _allreduce2d(POS, group)
                if not (kl and mask is None):
                    _allreduce2d(NEGD, group)
            ops.mu_update(D, POS, NEGD, Draw, skip=latch)
        # ---- l2_strict + max|D - D_new| < tol (batch_mu.py:21-23)
        ops.normalize_rows(Draw, Dn, False, True, D_ref=D if checks else None, tol=tol, latch=latch,
                           latch_value=it, maxdiff=maxdiff if checks else None, scratch=scratch, skip=latch)
        it_done = it
    if checks and not stopped_at:
        stopped_at = int(latch.item())
    if stopped_at:
        return stopped_at, Dbuf[stopped_at % 2], X
    if maxiter <= 1:
        return maxiter, Dbuf[0], X
    return maxiter, Dbuf[it_done % 2], X


def _allreduce2d(t, group):
    """Sum a (possibly row-padded) 2-D statistic over the ranks, in place."""
    if t.is_contiguous():
        torch.distributed.all_reduce(t, group=group)
    else:
        flat = t.contiguous()
        torch.distributed.all_reduce(flat, group=group)
        t.copy_(flat)
