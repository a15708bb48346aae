"""Minibatch multiplicative-update NMF drivers on the B200 ("next" row of the scope table).

Reference: decomp/nmf.py:82-111 -> decomp/nmf_methods/serizel.py (asg-mu, gsg-mu, asag-mu, gsag-mu) and
decomp/nmf_methods/kasai.py (svrmu, svrmu-acc).  The drivers are host loops over the same GEMM + fused-epilogue
kernels as the full-batch solver, applied to row blocks of the (device-side) shuffled data; the reference's quirks
are kept: 'gsg-mu' runs the 'asg-mu' loop (serizel.py:23-25) and on convergence the *previous* D is returned
(serizel.py:58, kasai.py:81).
"""
import numpy as np
import torch

from . import ops
from ._device import empty2d, full2d
from ._lib import rview

MINIBATCH_METHODS = ['asg-mu', 'gsg-mu', 'asag-mu', 'gsag-mu', 'svrmu', 'svrmu-acc']


class MuParts(object):
    """x update and the positive / negative parts of the D gradient for one row block (grads.py:77-160)."""

    def __init__(self, rows, f, k, kl, masked, dev):
        self.kl, self.masked, self.k, self.f = kl, masked, k, f
        self.Dt = empty2d(f, k, False, dev)
        self.NEG = empty2d(rows, k, False, dev)
        self.ws = ops.gemm_tn_workspace_for([(k, f, rows), (k, k, rows)], dev)
        if masked or kl:
            self.F = empty2d(rows, f, False, dev)
        if masked:
            self.YM = empty2d(rows, f, False, dev)
        if not masked and not kl:
            self.G = empty2d(k, k, False, dev)
            self.S = empty2d(k, k, False, dev)
        if kl and not masked:
            self.ones_kf = full2d(k, f, 1.0, False, dev)
            self.dsum = torch.empty(k + (k & 1), dtype=torch.float64, device=dev)[:k]
            self.xsum = torch.empty(k, dtype=torch.float64, device=dev)
        self.D = None

    def set_D(self, D):
        """Quantities that depend on D only (called whenever D changes)."""
        self.D = D
        ops.make_rhs(D, False, False, out=self.Dt)
        if not self.masked and not self.kl:
            ops.gemm_nt(D, D, ops.epilogue(ops.EPI_STORE, self.G))
        if self.kl and not self.masked:
            ops.row_sums(D, 1.0, out=self.dsum)

    def _ym(self, y, m):
        if not self.masked:
            return y
        ym = self.YM[:y.shape[0]]
        ops.mask_mul(y, m, ym)
        return ym

    def update_x(self, y, x, m):
        """x <- x * max(pos, 0) / max(neg, eps) in place (grads.py:77-84 with :108-115 / :142-149)."""
        E, D, Dt, r = ops.epilogue, self.D, self.Dt, y.shape[0]
        NEG = self.NEG[:r]
        if not self.kl:
            if not self.masked:
                ops.gemm_nt(x, self.G, E(ops.EPI_STORE, NEG))
                ops.gemm_nt(y, D, E(ops.EPI_MU_NUM, x, x=x, other=NEG))
            else:
                F = self.F[:r]
                ops.gemm_nt(x, Dt, E(ops.EPI_STORE_MASK, F, mask=m))
                ops.gemm_nt(F, D, E(ops.EPI_STORE, NEG))
                ops.gemm_nt(self._ym(y, m), D, E(ops.EPI_MU_NUM, x, x=x, other=NEG))
        else:
            F = self.F[:r]
            ops.gemm_nt(x, Dt, E(ops.EPI_KL_RATIO, F, other=y, mask=m))
            if not self.masked:
                e = E(ops.EPI_MU_NUM, x, x=x, other=self.dsum.view(1, self.k))
                e.ldother = 0
                ops.gemm_nt(F, D, e)
            else:
                ops.gemm_nt(m, D, E(ops.EPI_STORE, NEG))
                ops.gemm_nt(F, D, E(ops.EPI_MU_NUM, x, x=x, other=NEG))

    def grad_d(self, y, x, m, POS, NEGD):
        """POS, NEGD [k, f] <- positive / negative parts of the D gradient (grads.py:117-125 / :151-160)."""
        E, D, Dt, r, ws = ops.epilogue, self.D, self.Dt, y.shape[0], self.ws
        if not self.kl:
            if not self.masked:
                ops.gemm_tn(x, y, POS, workspace=ws)
                ops.gemm_tn(x, x, self.S, workspace=ws)
                ops.gemm_nt(self.S, Dt, E(ops.EPI_STORE, NEGD))
            else:
                F = self.F[:r]
                ops.gemm_nt(x, Dt, E(ops.EPI_STORE_MASK, F, mask=m))
                ops.gemm_tn(x, self._ym(y, m), POS, workspace=ws)
                ops.gemm_tn(x, F, NEGD, workspace=ws)
        else:
            F = self.F[:r]
            ops.gemm_nt(x, Dt, E(ops.EPI_KL_RATIO, F, other=y, mask=m))
            ops.gemm_tn(x, F, POS, workspace=ws)
            if not self.masked:
                ops.col_sums(x, 1.0, out=self.xsum)
                ops.scale(self.ones_kf, NEGD, rowscale=self.xsum)
            else:
                ops.gemm_tn(x, m, NEGD, workspace=ws)


class _ShuffledRows(object):
    """Device rows under the reference's cumulative shuffle (utils/data.py:124-156): two owned buffers are
    used alternately as gather targets; the caller's array is only ever read."""

    def __init__(self, array, cplx):
        self.cur = array
        self.cplx = cplx
        self.spare = [None, None]
        self.turn = 0

    def shuffle(self, index_dev):
        n, c = self.cur.shape
        if self.spare[self.turn] is None:
            self.spare[self.turn] = empty2d(n, c, self.cplx, self.cur.device)
        dst = self.spare[self.turn]
        ops.gather_rows(rview(self.cur), index_dev, rview(dst))
        self.cur = dst
        self.turn ^= 1

    def rows(self, r, step):
        return self.cur[r * step:(r + 1) * step]


def solve_device(y, D0, x, tol, minibatch, maxiter, method, kl, mask, rng, forget_rate=0.5, alpha=1.0, beta=0.5):
    """All six drivers on device tensors; returns ``(it, D, x)`` with x in the caller's row order."""
    dev = y.device
    n, f = y.shape
    k = D0.shape[0]
    masked = mask is not None
    parts = MuParts(minibatch, f, k, kl, masked, dev)
    ys, xs = _ShuffledRows(y, False), _ShuffledRows(x, False)
    ms = _ShuffledRows(mask, False) if masked else None
    index, restore = np.arange(n), np.arange(n)
    n_loop = n // minibatch

    def kf():
        return empty2d(k, f, False, dev)

    D, Dn, Draw, POS, NEGD = kf(), kf(), kf(), kf(), kf()
    ops.normalize_rows(D0, D, False, True)                                      # nmf.py:70
    result = torch.zeros(2, dtype=torch.float64, device=dev)
    scratch = torch.zeros(1, dtype=torch.int32, device=dev)

    def permute():
        nonlocal restore
        rng.shuffle(index)
        idx = torch.from_numpy(index).to(dev)
        ys.shuffle(idx)
        xs.shuffle(idx)
        if masked:
            ms.shuffle(idx)
        restore = restore[index]

    def batch(b):
        return ys.rows(b, minibatch), xs.rows(b, minibatch), (ms.rows(b, minibatch) if masked else None)

    def restored_x():
        order = torch.from_numpy(np.argsort(restore)).to(dev)
        out = empty2d(n, k, False, dev)
        ops.gather_rows(xs.cur, order, out)
        return out

    def converged():
        """normalise Draw into Dn and evaluate max|D - Dn| < tol (one small D2H read only when tol > 0)."""
        ops.normalize_rows(Draw, Dn, False, True)
        if tol > 0.0:
            ops.max_abs_diff(D, Dn, False, result, scratch)
            return float(result[1].item()) < tol
        return False

    if method in ('asg-mu', 'gsg-mu', 'asag-mu', 'gsag-mu'):
        accumulate = method in ('asag-mu', 'gsag-mu')
        every_batch = method != 'gsag-mu'
        if accumulate:
            PS, NS = kf(), kf()
        for it in range(1, maxiter):
            permute()
            if accumulate:
                PS.zero_()
                NS.zero_()
            for b in range(n_loop):
                y_mb, x_mb, m_mb = batch(b)
                parts.set_D(D)
                parts.update_x(y_mb, x_mb, m_mb)
                parts.grad_d(y_mb, x_mb, m_mb, POS, NEGD)
                if accumulate:                                                   # serizel.py:94-95
                    ops.axpby(1.0 - forget_rate, PS, forget_rate, POS, PS)
                    ops.axpby(1.0 - forget_rate, NS, forget_rate, NEGD, NS)
                if every_batch:
                    ops.mu_update(D, PS if accumulate else POS, NS if accumulate else NEGD, Draw)
                    if converged():
                        return it, D, restored_x()
                    D, Dn = Dn, D
            if not every_batch:
                ops.mu_update(D, PS, NS, Draw)
                if converged():
                    return it, D, restored_x()
                D, Dn = Dn, D
        return maxiter, D, restored_x()

    if method in ('svrmu', 'svrmu-acc'):
        inner = 1
        if method == 'svrmu-acc':                                                # kasai.py:23-27 (F, K = D.shape)
            F_, K_, N_ = k, f, n
            inner = int(np.maximum(beta * F_ * (3 * K_ + 2 * N_) / (3 * F_ * N_ + 2 * K_), 1.0))
        permute()                                                                # once (kasai.py:41-44)
        pos_prev = torch.zeros((n_loop, k, f), dtype=torch.float64, device=dev)
        neg_prev = torch.zeros((n_loop, k, f), dtype=torch.float64, device=dev)
        PF, NF, P, Q = kf(), kf(), kf(), kf()
        for it in range(1, maxiter):
            PF.zero_()
            NF.zero_()
            parts.set_D(D)
            for b in range(n_loop):                                              # full gradient (kasai.py:51-58)
                y_mb, x_mb, m_mb = batch(b)
                parts.grad_d(y_mb, x_mb, m_mb, POS, NEGD)
                ops.axpby(1.0, PF, 1.0, POS, PF)
                ops.axpby(1.0, NF, 1.0, NEGD, NF)
            ops.axpby(1.0 / n_loop, PF, 0.0, PF, PF)
            ops.axpby(1.0 / n_loop, NF, 0.0, NF, NF)
            for b in range(n_loop):
                y_mb, x_mb, m_mb = batch(b)
                parts.set_D(D)
                for _ in range(inner):
                    parts.update_x(y_mb, x_mb, m_mb)
                parts.grad_d(y_mb, x_mb, m_mb, POS, NEGD)
                ops.axpby(1.0, POS, 1.0, neg_prev[b], P)                          # kasai.py:72-73
                ops.axpby(1.0, P, 1.0, PF, P)
                ops.axpby(1.0, NEGD, 1.0, pos_prev[b], Q)
                ops.axpby(1.0, Q, 1.0, NF, Q)
                ops.svrmu_update(D, P, Q, alpha, Draw)
                if converged():
                    return it, D, restored_x()
                D, Dn = Dn, D
                pos_prev[b].copy_(POS)
                neg_prev[b].copy_(NEGD)
        return maxiter, D, restored_x()
    raise NotImplementedError('NMF with {} algorithm is not yet implemented.'.format(method))
