"""Online dictionary learning (Mairal et al. block coordinate descent) on the B200.

Drop-in for the reference's ``decomp.dictionary_learning.solve`` (decomp/dictionary_learning.py:12-231):
same signature, defaults, validation and return tuple ``(it, D, x)``; real or complex data, optional
missing-value mask.

    argmin_{x, D}  1/(2n) |y - x D|^2 + alpha |x|,   |D_k| <= 1        y [n, f], x [n, k], D [k, f]

Per minibatch, everything stays on the device:

    code      x_mb <- lasso(y_mb, D, alpha, x_mb; lasso_iter iterations)        decomp_b200.lasso.lasso_device
    stats     S <- beta S + x^H x   [k, k] ;  T <- beta T + x^H y   [k, f]      TN GEMM, contraction over the rows
              masked:  S[a] <- beta S[a] + mask^T (conj(x_a) x)  [k, f, k] ;  T <- beta T + x^H (y*mask)
    update    unmasked: Gauss-Seidel atom sweep, one cooperative launch         dictionary_learning.py:154-159
              masked:   Jacobi update streaming S once                          dictionary_learning.py:216-222
    test      max|D - D_new| < tol                                               dictionary_learning.py:161,224

The epoch shuffle uses ``numpy.random.RandomState(random_seed)`` on the host exactly like the reference
(the permutation is cumulative, utils/data.py:147-156); the rows are permuted on the device by a gather kernel
and the final ``x`` is un-permuted the same way.
"""
import numpy as np
import torch

from . import ops
from ._device import array_kind, empty2d, is_torch, np_dtype, require_cuda, to_device2d, to_host, zeros2d
from ._lib import rview
from .lasso import DEVICE_RULES, AVAILABLE_METHODS, lasso_device
from .utils import assertion


def solve(y, D, alpha, x=None, tol=1.0e-3, minibatch=None, maxiter=1000, method='block_cd', lasso_method='cd',
          lasso_iter=10, lasso_tol=1.0e-5, mask=None, random_seed=None, group=None):
    """Learn the dictionary ``D`` and the codes ``x``; see the module docstring.

    ``group`` (not in the reference): optional ``torch.distributed`` process group.  The algorithm is sequential
    across minibatches, so the ranks share each minibatch: every rank passes the SAME arrays and seed, codes its
    own block of rows of each minibatch, the new codes are exchanged, the sufficient statistics are all-reduced and
    the atom update is replicated.  Every rank returns the same ``(it, D, x)``.

    ``lasso_method`` must be one of the device rules ('ista', 'fista', 'acc_ista', optionally '_pos'); the
    reference's default 'cd' is a sequential reference-purpose method outside the hot path.
    """
    array_kind(y, D, x, mask)
    if x is None:
        if is_torch(D):
            x = torch.ones((y.shape[0], D.shape[0]), dtype=D.dtype, device=D.device)
        else:
            x = np.ones((y.shape[0], D.shape[0]), dtype=D.dtype)

    assertion.assert_dtypes(y=y, D=D, x=x)
    assertion.assert_dtypes(mask=mask, dtypes='f')
    assertion.assert_shapes('x', x, 'D', D, axes=1)
    assertion.assert_shapes('y', y, 'D', D, axes=[-1])
    assertion.assert_shapes('y', y, 'mask', mask)

    if minibatch is None:
        raise NotImplementedError('Only online methods are implemented. minibatch is required.')
    if y.shape[0] < minibatch:
        raise ValueError('Minibatch size should be smaller than the total size. Given {} < {}'.format(
            y.shape[0], minibatch))
    if method != 'block_cd':
        raise NotImplementedError('Method %s is not yet implemented' % method)
    positive = lasso_method[-4:] == '_pos'
    rule = lasso_method[:-4] if positive else lasso_method
    if rule not in DEVICE_RULES:
        if rule in AVAILABLE_METHODS:
            raise NotImplementedError("lasso_method '%s' is not on the B200 hot path; use 'ista', 'fista' or "
                                      "'acc_ista'." % lasso_method)
        raise NotImplementedError('Method ' + lasso_method + ' is not yet implemented.')

    device = require_cuda()
    out_dtype = np_dtype(y)
    yd = to_device2d(y, device, copy=False)
    md = to_device2d(mask, device, copy=False) if mask is not None else None
    Dd = to_device2d(D, device, copy=True)
    xd = to_device2d(x, device, copy=True)
    rng = np.random.RandomState(random_seed)
    it, Dd, xd = block_cd_device(yd, Dd, float(alpha), xd, float(tol), int(minibatch), int(maxiter), rule, positive,
                                 int(lasso_iter), float(lasso_tol), md, rng, group=group)
    return it, to_host(Dd, y, out_dtype), to_host(xd, y, out_dtype)


def _pair_cols(f, cw, device):
    """(atom, b) pairs per statistics GEMM: the [f, pairs*cw] output is cut into 128x64 tiles that a persistent grid
    of one CTA per SM walks in rounds, so pick the width whose tile count fills whole rounds best."""
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    tiles_m = (f + 127) // 128
    best, best_eff = 8, 0.0
    for tiles_n in range(8, 65):
        tiles = tiles_m * tiles_n
        eff = tiles / float(-(-tiles // sms) * sms)
        if eff > best_eff + 1e-9:
            best, best_eff = tiles_n, eff
    return best * 64 // cw


def _pair_chunks(k, cap, device):
    """[(colA, colB)] int32 device vectors listing the pairs (a, b >= a) in row-major order, ~cap pairs each."""
    a_idx, b_idx = np.triu_indices(k)
    out = []
    for s0 in range(0, a_idx.size, cap):
        out.append((torch.from_numpy(a_idx[s0:s0 + cap].astype(np.int32)).to(device),
                    torch.from_numpy(b_idx[s0:s0 + cap].astype(np.int32)).to(device)))
    return out


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


def block_cd_device(y, D0, alpha, x, tol, minibatch, maxiter, rule, positive, lasso_iter, lasso_tol, mask, rng,
                    group=None):
    """``solve_cd`` / ``solve_cd_mask`` on device tensors. Returns ``(it, D, x)`` with x in the caller's row order."""
    dev = y.device
    n, f = y.shape
    k = D0.shape[0]
    cplx = y.is_complex()
    cw = 2 if cplx else 1
    masked = mask is not None
    stat_combine = 3 if cplx else 1

    ys, xs = _ShuffledRows(y, cplx), _ShuffledRows(x, x.is_complex())
    ms = _ShuffledRows(mask, False) if masked else None
    index = np.arange(n)
    restore = np.arange(n)

    D = empty2d(k, f, cplx, dev)
    Dn = empty2d(k, f, cplx, dev)
    ops.normalize_rows(rview(D0), rview(D), cplx, True)                       # :125, :182
    T = zeros2d(k, f, cplx, dev)
    if masked:
        S = torch.zeros((k, f, k * cw), dtype=torch.float64, device=dev)      # [k][f][k] (interleaved complex)
        YM = empty2d(minibatch, f, cplx, dev)
        Dt_ws = torch.empty(f * k * cw, dtype=torch.float64, device=dev)
        # the (atom a, b >= a) pairs of the Hermitian half of S, packed into wide GEMMs
        chunks = _pair_chunks(k, _pair_cols(f, cw, dev), dev)
        widest = max(c[0].numel() for c in chunks)
        Wt = empty2d(widest * cw, minibatch, False, dev)     # transposed pair products: contraction index contiguous
        Xt = empty2d(k * cw, minibatch, False, dev)          # transposed codes and mask of the minibatch
        Mt = empty2d(f, minibatch, False, dev)
        Ptmp = empty2d(f, widest, cplx, dev)
        ws = ops.gemm_tn_workspace_for([(k * cw, f * cw, minibatch)], dev)
    else:
        S = zeros2d(k, k, cplx, dev)
        ws = ops.gemm_tn_workspace_for([(k * cw, k * cw, minibatch), (k * cw, f * cw, minibatch)], dev)
        sweep_ws = ops.dl_sweep_workspace(k, f, cplx, dev)
    dist = torch.distributed if group is not None else None
    if dist is not None:
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        lo, hi = rank * minibatch // world, (rank + 1) * minibatch // world     # this rank's rows of a minibatch
        S_part = torch.zeros_like(S)            # local statistics before the all-reduce
        T_part = zeros2d(k, f, cplx, dev)
    else:
        lo, hi = 0, minibatch
    result = torch.zeros(2, dtype=torch.float64, device=dev)
    scratch = torch.zeros(1, dtype=torch.int32, device=dev)
    checks = tol > 0.0

    def restored_x():
        order = torch.from_numpy(np.argsort(restore)).to(dev)
        out = empty2d(n, k, xs.cplx, dev)
        ops.gather_rows(rview(xs.cur), order, rview(out))
        return out

    count = 0
    for it in range(1, maxiter):
        rng.shuffle(index)                                                     # :131-133 (cumulative)
        index_dev = torch.from_numpy(index).to(dev)
        ys.shuffle(index_dev)
        xs.shuffle(index_dev)
        if masked:
            ms.shuffle(index_dev)
        restore = restore[index]
        try:
            for r in range(n // minibatch):                                    # tail rows are skipped
                y_mb, x_mb = ys.rows(r, minibatch), xs.rows(r, minibatch)
                m_mb = ms.rows(r, minibatch) if masked else None
                if dist is None:
                    lasso_device(y_mb, D, alpha, x_mb, lasso_tol, lasso_iter, rule, positive, m_mb, out=x_mb)
                else:
                    # code this rank's rows, then make the whole minibatch's codes known everywhere
                    x_own = x_mb[lo:hi]
                    lasso_device(y_mb[lo:hi], D, alpha, x_own, lasso_tol, lasso_iter, rule, positive,
                                 m_mb[lo:hi] if masked else None, out=x_own, group=group)
                    x_mb[:lo].zero_()
                    x_mb[hi:].zero_()
                    _allreduce(x_mb, group)

                theta = count * minibatch + 1.0                                # equation (11), :143-144
                beta = (theta - minibatch) / theta
                # statistics over this rank's rows (all rows without a group); with a group the partial sums are
                # all-reduced and folded in as  S <- beta S + sum_ranks(partial)
                xr = rview(x_mb[lo:hi])
                rows = hi - lo
                S_dst, T_dst = (S, T) if dist is None else (S_part, T_part)
                comb = stat_combine if dist is None else stat_combine - 1       # accumulate / overwrite
                if not masked:
                    ops.gemm_tn(xr, xr, rview(S_dst), combine=comb, beta=beta, workspace=ws)           # :151
                    ops.gemm_tn(xr, rview(y_mb[lo:hi]), rview(T_dst), combine=comb, beta=beta, workspace=ws)
                else:
                    # S[a][j][b] = sum_i conj(x_ia) x_ib m_ij is Hermitian in (a, b): accumulate b >= a, mirror the rest
                    # as NT GEMMs  mask^T [f, rows] . Wt [pairs, rows]^T  (the NT kernel is the faster one)
                    ops.make_rhs(xr, False, False, out=Xt[:, :rows])
                    ops.make_rhs(m_mb[lo:hi], False, False, out=Mt[:, :rows])
                    for colA, colB in chunks:                                                          # :210-213
                        wd = colA.numel()
                        Wc, Pc = Wt[:wd * cw, :rows], rview(Ptmp[:, :wd])
                        ops.dl_pair_products_t(Xt[:, :rows], cplx, colA, colB, Wc)
                        ops.gemm_nt(Mt[:, :rows], Wc, ops.epilogue(ops.EPI_STORE, Pc))
                        ops.dl_scatter_stats(Pc, cplx, colA, colB, k, beta if dist is None else 0.0, S_dst)
                    ops.dl_mirror(S_dst, k, f, cplx)
                    ops.mask_mul(rview(y_mb[lo:hi]), m_mb[lo:hi], rview(YM[:rows]), cwidth=cw)
                    ops.gemm_tn(xr, rview(YM[:rows]), rview(T_dst), combine=comb, beta=beta, workspace=ws)  # :214
                if dist is not None:
                    _allreduce(S_part, group)
                    _allreduce(T_part, group)
                    S2, P2 = (S.view(k * f, k * cw), S_part.view(k * f, k * cw)) if masked else (rview(S), rview(S_part))
                    ops.axpby(beta, S2, 1.0, P2, S2)
                    ops.axpby(beta, rview(T), 1.0, rview(T_part), rview(T))
                if not masked:
                    Dn.copy_(D)
                    ops.dl_sweep(rview(S), rview(T), rview(Dn), cplx, ws=sweep_ws)                            # :154-159
                else:
                    ops.dl_masked_update(S, rview(T), rview(D), rview(Dn), cplx, Dt_ws)                # :216-222
                if checks:
                    ops.max_abs_diff(rview(D), rview(Dn), cplx, result, scratch)
                    if float(result[1].item()) < tol:                                                  # :161, :224
                        return it, Dn, restored_x()
                D, Dn = Dn, D
                count += 1
        except KeyboardInterrupt:
            return it, D, restored_x()
    return maxiter, D, restored_x()


def _allreduce(t, group):
    """Sum over the ranks, in place, also for row-padded (non-contiguous) buffers."""
    if t.is_contiguous():
        torch.distributed.all_reduce(t, group=group)
    else:
        flat = t.contiguous()
        torch.distributed.all_reduce(flat, group=group)
        t.copy_(flat)
