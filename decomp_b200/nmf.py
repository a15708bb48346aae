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

    if likelihood in ('l2', 'gaussian'):
        kl = False
    elif likelihood in ('kl', 'poisson'):
        kl = True
    else:
        raise NotImplementedError('Likelihood {} is not implemented for nmf'.format(likelihood))
    if minibatch is not None:
        return _solve_minibatch(y, D, x, tol, minibatch, maxiter, method, kl, mask, random_seed, group, kwargs)
    if method != 'mu':
        raise NotImplementedError('Batch-NMF with {} algorithm is not yet implemented.'.format(method))
    if kwargs:
        raise TypeError('solve() got unexpected keyword arguments ' + str(sorted(kwargs)))

    device = require_cuda()
    out_dtype = np_dtype(y)
    yd = to_device2d(y, device, copy=False)
    md = to_device2d(mask, device, copy=False) if mask is not None else None
    Dd = to_device2d(D, device, copy=True)
    xd = to_device2d(x, device, copy=True)       # updated in place: always our own copy
    it, Dd, xd = mu_device(yd, Dd, xd, float(tol), int(maxiter), kl, md, group=group)
    return it, to_host(Dd, y, out_dtype), to_host(xd, y, out_dtype)


def _solve_minibatch(y, D, x, tol, minibatch, maxiter, method, kl, mask, random_seed, group, kwargs):
    """decomp/nmf.py:82-111: the stochastic drivers (serizel.py, kasai.py) on row blocks of shuffled data."""
    from . import nmf_minibatch
    if method not in MINIBATCH_METHODS:
        raise NotImplementedError('NMF with {} algorithm is not yet implemented.'.format(method))
    if group is not None:
        raise NotImplementedError('minibatch NMF runs on one GPU')
    allowed = ('forget_rate',) if method.endswith('g-mu') else ('alpha', 'beta')
    for key in kwargs:
        if key not in allowed:
            raise TypeError("solve() got an unexpected keyword argument '%s'" % key)
    if y.shape[0] < minibatch:                                        # utils/data.py:79-82
        raise ValueError('Minibatch size should be smaller than the total size. Given {} < {}'.format(
            y.shape[0], minibatch))
    device = require_cuda()
    out_dtype = np_dtype(y)
    yd = to_device2d(y, device, copy=False)
    md = to_device2d(mask, device, copy=False) if mask is not None else None
    Dd = to_device2d(D, device, copy=True)
    xd = to_device2d(x, device, copy=True)
    rng = np.random.RandomState(random_seed)
    it, Dd, xd = nmf_minibatch.solve_device(yd, Dd, xd, float(tol), int(minibatch), int(maxiter), method, kl, md, rng,
                                            **kwargs)
    return it, to_host(Dd, y, out_dtype), to_host(xd, y, out_dtype)


def mu_device(y, D0, X, tol, maxiter, kl=False, mask=None, group=None):
    """Full-batch MU on device tensors; ``X`` [n, k] is updated in place. Returns ``(it, D, X)``."""
    solver = MuSolver(y, D0, X, tol, kl=kl, mask=mask, group=group)
    stopped_at = 0
    for it in range(1, maxiter):
        if solver.checks and it % POLL_EVERY == 0:
            stopped_at = solver.fired()
            if stopped_at:
                break
        solver.sweep(it)
    if solver.checks and not stopped_at:
        stopped_at = solver.fired()
    if stopped_at:
        return stopped_at, solver.Dbuf[stopped_at % 2], X
    return maxiter, solver.Dbuf[max(maxiter - 1, 0) % 2], X


class MuSolver(object):
    """Device state of a full-batch MU run; ``sweep(it)`` enqueues sweep number ``it`` (1-based, reading
    ``Dbuf[(it - 1) % 2]`` and writing ``Dbuf[it % 2]``) on the current stream without synchronising."""

    def __init__(self, y, D0, X, tol, kl=False, mask=None, group=None):
        dev = y.device
        self.y, self.X, self.mask, self.kl, self.tol, self.group = y, X, mask, kl, tol, group
        self.n, self.f = n, f = y.shape
        self.k = k = D0.shape[0]
        self.Dbuf = [empty2d(k, f, False, dev), empty2d(k, f, False, dev)]
        ops.normalize_rows(D0, self.Dbuf[0], False, True)                 # nmf.py:70
        self.Draw = empty2d(k, f, False, dev)
        self.Dt = empty2d(f, k, False, dev)
        self.POS = empty2d(k, f, False, dev)
        self.NEG = empty2d(n, k, False, dev)
        self.ws = ops.gemm_tn_workspace_for([(k, f, n), (k, k, n)], dev)
        self.checks = tol > 0.0
        self.latch = torch.zeros(1, dtype=torch.int32, device=dev) if self.checks else None
        self.scratch = torch.zeros(1, dtype=torch.int32, device=dev)
        self.maxdiff = torch.zeros(2, dtype=torch.float64, device=dev)
        self.ym = y
        if mask is not None:
            self.ym = empty2d(n, f, False, dev)
            ops.mask_mul(y, mask, self.ym)                                # y * mask, once (grads.py:113,123)
        if mask is not None or kl:
            self.F = empty2d(n, f, False, dev)                            # the [n, f] intermediate
            self.NEGD = empty2d(k, f, False, dev)
        else:
            self.G = empty2d(k, k, False, dev)
            self.S = empty2d(k, k, False, dev)
        if kl and mask is None:
            self.ones_kf = full2d(k, f, 1.0, False, dev)
            self.dsum = torch.empty(k + (k & 1), dtype=torch.float64, device=dev)[:k]   # even: 16-byte epilogue loads
            self.xsum = torch.empty(k, dtype=torch.float64, device=dev)

    def fired(self):
        return int(self.latch.item()) if self.latch is not None else 0

    def sweep(self, it):
        E = ops.epilogue
        y, ym, X, mask, latch, ws, group = self.y, self.ym, self.X, self.mask, self.latch, self.ws, self.group
        D, Dn = self.Dbuf[(it - 1) % 2], self.Dbuf[it % 2]
        Dt, POS, NEG, Draw = self.Dt, self.POS, self.NEG, self.Draw
        if not self.kl and mask is None:
            G, S = self.G, self.S
            # ---- x update (grads.py:108-111 with f.dot(d.T) re-associated)
            ops.gemm_nt(D, D, E(ops.EPI_STORE, G), skip=latch)
            ops.gemm_nt(X, G, E(ops.EPI_STORE, NEG), skip=latch)
            ops.gemm_nt(y, D, E(ops.EPI_MU_NUM, X, x=X, other=NEG), skip=latch)
            # ---- D update (grads.py:117-121): sufficient statistics over the sample axis
            ops.gemm_tn(X, y, POS, workspace=ws, skip=latch)
            ops.gemm_tn(X, X, S, workspace=ws, skip=latch)
            if group is not None:
                _allreduce2d(POS, group)
                _allreduce2d(S, group)
            ops.make_rhs(D, False, False, out=Dt, skip=latch)
            ops.gemm_nt(S, Dt, E(ops.EPI_MU_DEN, Draw, x=D, other=POS), skip=latch)
        else:
            F, NEGD = self.F, self.NEGD
            ops.make_rhs(D, False, False, out=Dt, skip=latch)
            if not self.kl:
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
                    ops.row_sums(D, 1.0, out=self.dsum)
                    neg_x = E(ops.EPI_MU_NUM, X, x=X, other=self.dsum.view(1, self.k))
                    neg_x.ldother = 0                                  # one [1, k] row for every sample
                    ops.gemm_nt(F, D, neg_x, skip=latch)
                else:
                    ops.gemm_nt(mask, D, E(ops.EPI_STORE, NEG), skip=latch)
                    ops.gemm_nt(F, D, E(ops.EPI_MU_NUM, X, x=X, other=NEG), skip=latch)
                ops.gemm_nt(X, Dt, E(ops.EPI_KL_RATIO, F, other=y, mask=mask), skip=latch)
                ops.gemm_tn(X, F, POS, workspace=ws, skip=latch)
                if mask is None:
                    ops.col_sums(X, 1.0, out=self.xsum)
                    if group is not None:
                        torch.distributed.all_reduce(self.xsum, group=group)
                    ops.scale(self.ones_kf, NEGD, rowscale=self.xsum)
                else:
                    ops.gemm_tn(X, mask, NEGD, workspace=ws, skip=latch)
            if group is not None:
                _allreduce2d(POS, group)
                if not (self.kl and mask is None):
                    _allreduce2d(NEGD, group)
            ops.mu_update(D, POS, NEGD, Draw, skip=latch)
        # ---- l2_strict + max|D - D_new| < tol (batch_mu.py:21-23)
        ops.normalize_rows(Draw, Dn, False, True, D_ref=D if self.checks else None, tol=self.tol, latch=latch,
                           latch_value=it, maxdiff=self.maxdiff if self.checks else None, scratch=self.scratch,
                           skip=latch)


def _allreduce2d(t, group):
    """Sum a (possibly row-padded) 2-D statistic over the ranks, in place."""
    if t.is_contiguous():
        torch.distributed.all_reduce(t, group=group)
    else:
        flat = t.contiguous()
        torch.distributed.all_reduce(flat, group=group)
        t.copy_(flat)
