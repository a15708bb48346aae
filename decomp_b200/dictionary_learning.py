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
import os

import numpy as np
import torch

from . import comm, ops
from ._device import array_kind, empty2d, is_torch, np_dtype, require_cuda, to_device2d, to_host, zeros2d
from ._lib import rview
from .lasso import DEVICE_RULES, AVAILABLE_METHODS, lasso_device
from .utils import assertion


def solve(y, D, alpha, x=None, tol=1.0e-3, minibatch=None, maxiter=1000, method='block_cd', lasso_method='cd',
          lasso_iter=10, lasso_tol=1.0e-5, mask=None, random_seed=None, group=None):
    """Learn the dictionary ``D`` and the codes ``x``; see the module docstring.

    ``group`` (not in the reference): optional ``torch.distributed`` process group.  Every rank passes its own
    contiguous block of rows of ``y`` / ``x`` / ``mask`` (rank order = row order), the same ``D`` and the same seed.
    The algorithm is sequential across minibatches, so the ranks share each minibatch: a rank codes the rows of the
    minibatch it owns; the statistics are all-reduced (masked: reduce-scattered along the feature axis, each rank
    updating its own channels, then the new dictionary is all-gathered).  Every rank returns the same ``it`` and
    ``D`` and its own rows of ``x``; the results are those of the single-process run on the concatenated rows.

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
    if group is None and y.shape[0] < minibatch:        # (sharded: checked against the total row count below)
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


OVERLAP_STATS_EXCHANGE = os.environ.get('DECOMP_DL_OVERLAP', '1') != '0'   # sharded masked statistics: reduce-scatter on
                                                                           # a side stream under the next chunk's GEMM


def _row_layout(n_local, group, device):
    """(total rows, first global row of this rank, ranks, rank) for row-sharded inputs."""
    if group is None:
        return n_local, 0, 1, 0
    dist = torch.distributed
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    counts[rank] = n_local
    dist.all_reduce(counts, group=group)
    counts = counts.cpu().numpy()
    return int(counts.sum()), int(counts[:rank].sum()), world, rank


def _epoch_selection(order, steps, minibatch, row0, n_loc):
    """The rows of each minibatch of one epoch that live on this rank (global rows row0 .. row0 + n_loc), as local row
    numbers in minibatch order: (concatenated int64 selections, bounds with bounds[r] .. bounds[r+1] = minibatch r).
    Over the ranks the selections of a minibatch partition ``order[r*mb:(r+1)*mb]``."""
    sel, bounds = [], [0]
    for r in range(steps):
        rows = order[r * minibatch:(r + 1) * minibatch]
        rows = rows[(rows >= row0) & (rows < row0 + n_loc)] - row0
        sel.append(rows)
        bounds.append(bounds[-1] + rows.size)
    cat = np.concatenate(sel).astype(np.int64) if sel else np.zeros(0, dtype=np.int64)
    return cat, bounds


def block_cd_device(y, D0, alpha, x, tol, minibatch, maxiter, rule, positive, lasso_iter, lasso_tol, mask, rng,
                    group=None):
    """``solve_cd`` / ``solve_cd_mask`` on device tensors. Returns ``(it, D, x)`` with x in the caller's row order.

    The data never moves: ``order`` (host) is the reference's cumulative permutation (utils/data.py:147-156,
    dictionary_learning.py:131-133) and minibatch r is the rows ``order[r*mb:(r+1)*mb]``, gathered into a work buffer,
    coded, and scattered back.  With a ``group`` every rank holds a contiguous block of rows and takes, from each
    minibatch, the rows it owns (their codes are row-local given D, lasso.py:244-246); what crosses the ranks per
    minibatch is: unmasked, the all-reduce of x^H x [k,k] and x^H y [k,f], the atom sweep being
    replicated; masked, the reduce-scatter of the [k,f,k] statistic ALONG f -- the update is channel-local
    (dictionary_learning.py:218), every rank keeps and updates f / ranks channels -- two [k] all-reduces
    (sum_j S[a][j][a], |u_a|^2) and the all-gather of the new dictionary."""
    dev = y.device
    n_loc, f = y.shape
    k = D0.shape[0]
    cplx = y.is_complex()
    cw = 2 if cplx else 1
    masked = mask is not None
    stat_combine = 3 if cplx else 1
    dist = torch.distributed if group is not None else None
    n, row0, world, rank = _row_layout(n_loc, group, dev)
    if n < minibatch:
        raise ValueError('Minibatch size should be smaller than the total size. Given {} < {}'.format(n, minibatch))
    sharded = world > 1

    order = np.arange(n)
    index = np.arange(n)
    D = empty2d(k, f, cplx, dev)
    Dn = empty2d(k, f, cplx, dev)
    ops.normalize_rows(rview(D0), rview(D), cplx, True)                       # :125, :182
    T = zeros2d(k, f, cplx, dev)
    # work buffers of one minibatch (this rank's rows of it: at most all of them)
    Ymb, Xmb = empty2d(minibatch, f, cplx, dev), empty2d(minibatch, k, x.is_complex(), dev)
    Mmb = empty2d(minibatch, f, False, dev) if masked else None
    if masked:
        fs = -(-f // world)                       # channels per rank (the last ranks may own fewer, or none)
        j0 = rank * fs
        S = torch.zeros((k, fs, k * cw), dtype=torch.float64, device=dev)     # this rank's channel slab [k][fs][k]
        YM = empty2d(minibatch, f, cplx, dev)
        Dt_ws = torch.empty(fs * k * cw, dtype=torch.float64, device=dev)
        # the (atom a, b >= a) pairs of the Hermitian half of S, packed into wide GEMMs
        chunks = _pair_chunks(k, _pair_cols(f, cw, dev), dev)
        widest = max(c[0].numel() for c in chunks)
        Wt = empty2d(widest * cw, minibatch, False, dev)     # transposed pair products: contraction index contiguous
        Xt = empty2d(k * cw, minibatch, False, dev)          # transposed codes and mask of the minibatch
        Mt = empty2d(f, minibatch, False, dev)
        ws = ops.gemm_tn_workspace_for([(k * cw, f * cw, minibatch)], dev)
        if not sharded:
            Ptmp = empty2d(f, widest, cplx, dev)
        else:
            # the packed pair statistics [f, pairs] of a chunk are reduce-scattered ALONG f as they are (only the
            # Hermitian half crosses the links, and no dense [k, f, k] send buffer exists): contiguous send buffers
            # of fs * world rows (rows beyond f stay zero), one per distinct chunk width
            # (row pitch rounded up to an even number of doubles: the GEMM epilogue stores 16-byte pairs)
            # Two buffers per width: the reduce-scatter of chunk i runs on a side stream while the pair-product GEMM of
            # chunk i + 1 fills the other buffer on the compute stream (OVERLAP_STATS_EXCHANGE)
            Psend = {wd: [torch.zeros((fs * world, wd * cw + (wd * cw & 1)), dtype=torch.float64, device=dev)
                          for _ in range(2)] for wd in set(c[0].numel() for c in chunks)}
            Precv = {wd: [torch.empty((fs, wd * cw + (wd * cw & 1)), dtype=torch.float64, device=dev)
                          for _ in range(2)] for wd in Psend}
            comm_stream = torch.cuda.Stream(device=dev) if OVERLAP_STATS_EXCHANGE else None
            ev_filled = [torch.cuda.Event() for _ in range(2)]       # send buffer written (compute stream)
            ev_reduced = [torch.cuda.Event() for _ in range(2)]      # receive buffer complete (side stream)
            ev_folded = [torch.cuda.Event() for _ in range(2)]       # receive buffer consumed (compute stream)
            stats = torch.zeros(k * 4, dtype=torch.float64, device=dev)
            D_slab = torch.zeros((k, fs * cw), dtype=torch.float64, device=dev)
            D_all = torch.empty((world, k, fs * cw), dtype=torch.float64, device=dev)
    else:
        S = zeros2d(k, k, cplx, dev)
        ws = ops.gemm_tn_workspace_for([(k * cw, k * cw, minibatch), (k * cw, f * cw, minibatch)], dev)
        sweep_ws = ops.dl_sweep_workspace(k, f, cplx, dev)
    if sharded:
        S_loc = zeros2d(k, k, cplx, dev) if not masked else None    # local statistics before the all-reduce
        T_part = zeros2d(k, f, cplx, dev)
    result = torch.zeros(2, dtype=torch.float64, device=dev)
    scratch = torch.zeros(2, dtype=torch.int32, device=dev)
    checks = tol > 0.0
    steps = n // minibatch                                                      # tail rows are skipped

    count = 0
    for it in range(1, maxiter):
        rng.shuffle(index)                                                     # :131-133 (cumulative)
        order = order[index]
        # this rank's rows of every minibatch of the epoch, as local row numbers: one upload per epoch
        sel, bounds = _epoch_selection(order, steps, minibatch, row0, n_loc)
        sel_dev = torch.from_numpy(sel).to(dev) if sel.size else None
        try:
            for r in range(steps):
                m = bounds[r + 1] - bounds[r]
                rows_dev = sel_dev[bounds[r]:bounds[r + 1]] if m else None
                y_mb, x_mb = Ymb[:m], Xmb[:m]
                m_mb = Mmb[:m] if masked else None
                if m:
                    ops.gather_rows(rview(y), rows_dev, rview(y_mb))
                    ops.gather_rows(rview(x), rows_dev, rview(x_mb))
                    if masked:
                        ops.gather_rows(mask, rows_dev, m_mb)
                lasso_device(y_mb, D, alpha, x_mb, lasso_tol, lasso_iter, rule, positive, m_mb, out=x_mb, group=group)
                if m:
                    ops.scatter_rows(rview(x_mb), rows_dev, rview(x))

                theta = count * minibatch + 1.0                                # equation (11), :143-144
                beta = (theta - minibatch) / theta
                xr = rview(x_mb)
                if not masked:
                    # ---- statistics (:147-152); sharded: local sums, all-reduced, folded in as S <- beta S + sum
                    S_dst, T_dst = (S, T) if not sharded else (S_loc, T_part)
                    comb = stat_combine if not sharded else stat_combine - 1    # accumulate / overwrite
                    if m:
                        ops.gemm_tn(xr, xr, rview(S_dst), combine=comb, beta=beta, workspace=ws)
                        ops.gemm_tn(xr, rview(y_mb), rview(T_dst), combine=comb, beta=beta, workspace=ws)
                    elif sharded:
                        S_dst.zero_()
                        T_dst.zero_()
                    if sharded:
                        _allreduce(S_loc, group)
                        _allreduce(T_part, group)
                        ops.axpby(beta, rview(S), 1.0, rview(S_loc), rview(S))
                        ops.axpby(beta, rview(T), 1.0, rview(T_part), rview(T))
                    Dn.copy_(D)
                    ops.dl_sweep(rview(S), rview(T), rview(Dn), cplx, ws=sweep_ws)                     # :154-159
                else:
                    # ---- S[a][j][b] = sum_i conj(x_ia) x_ib m_ij (:210-213) is Hermitian in (a, b): accumulate b >= a as
                    # NT GEMMs  mask^T [f, rows] . Wt [pairs, rows]^T, mirror the rest.  Sharded: every rank forms its
                    # rows' contribution to ALL channels of a chunk of pairs; the chunk is reduce-scattered along f and
                    # each rank folds its channels into its slab of S
                    T_dst = T if not sharded else T_part
                    comb = stat_combine if not sharded else stat_combine - 1
                    if m:
                        ops.make_rhs(xr, False, False, out=Xt[:, :m])
                        ops.make_rhs(m_mb, False, False, out=Mt[:, :m])
                    cur = torch.cuda.current_stream(dev)
                    pending = None                        # (slot, width, colA, colB) whose reduce-scatter is in flight
                    for ci, (colA, colB) in enumerate(chunks):
                        wd = colA.numel()
                        slot = ci & 1
                        Pc = rview(Ptmp[:, :wd]) if not sharded else Psend[wd][slot][:f, :wd * cw]
                        # (stream order already guarantees that the reduce-scatter which last read this send buffer,
                        # chunk ci - 2, is complete: its result was folded into S on this stream one chunk ago)
                        if m:
                            Wc = Wt[:wd * cw, :m]
                            ops.dl_pair_products_t(Xt[:, :m], cplx, colA, colB, Wc)
                            ops.gemm_nt(Mt[:, :m], Wc, ops.epilogue(ops.EPI_STORE, Pc))
                        else:
                            Pc.zero_()                    # (only a sharded run can leave a rank without rows)
                        if not sharded:
                            ops.dl_scatter_stats(Pc, cplx, colA, colB, k, beta, S)                      # S <- beta S + .
                            continue
                        if comm_stream is None:
                            comm.reduce_scatter_sum(Precv[wd][slot], Psend[wd][slot], group)            # along f
                            ops.dl_scatter_stats(Precv[wd][slot][:, :wd * cw], cplx, colA, colB, k, beta, S)
                            continue
                        ev_filled[slot].record(cur)
                        with torch.cuda.stream(comm_stream):
                            comm_stream.wait_event(ev_filled[slot])
                            comm_stream.wait_event(ev_folded[slot])   # the receive buffer's previous contents are used up
                            comm.reduce_scatter_sum(Precv[wd][slot], Psend[wd][slot], group)
                            ev_reduced[slot].record(comm_stream)
                        if pending is not None:           # fold the previous chunk in while this one is on the links
                            ps, pw, pA, pB = pending
                            cur.wait_event(ev_reduced[ps])
                            ops.dl_scatter_stats(Precv[pw][ps][:, :pw * cw], cplx, pA, pB, k, beta, S)
                            ev_folded[ps].record(cur)
                        pending = (slot, wd, colA, colB)
                    if pending is not None:
                        ps, pw, pA, pB = pending
                        cur.wait_event(ev_reduced[ps])
                        ops.dl_scatter_stats(Precv[pw][ps][:, :pw * cw], cplx, pA, pB, k, beta, S)
                        ev_folded[ps].record(cur)
                    if m:
                        ops.mask_mul(rview(y_mb), m_mb, rview(YM[:m]), cwidth=cw)
                        ops.gemm_tn(xr, rview(YM[:m]), rview(T_dst), combine=comb, beta=beta, workspace=ws)   # :214
                    elif sharded:
                        T_part.zero_()
                    if not sharded:
                        ops.dl_mirror(S, k, f, cplx)
                        ops.dl_masked_update(S, rview(T), rview(D), rview(Dn), cplx, Dt_ws)            # :216-222
                    else:
                        ops.dl_mirror(S, k, fs, cplx)
                        _allreduce(T_part, group)
                        ops.axpby(beta, rview(T), 1.0, rview(T_part), rview(T))
                        # channel-local Jacobi update of this rank's slab, two [k] sums over all channels exchanged
                        ops.dl_masked_update_phase(1, S, j0, rview(T), rview(D), cplx, D_slab, stats, Dt_ws)
                        comm.all_reduce_sum(stats, group)
                        ops.dl_masked_update_phase(2, S, j0, rview(T), rview(D), cplx, D_slab, stats, Dt_ws)
                        comm.all_reduce_sum(stats, group)
                        ops.dl_masked_update_phase(3, S, j0, rview(T), rview(D), cplx, D_slab, stats, Dt_ws)
                        comm.all_gather(D_all, D_slab, group)
                        rview(Dn).copy_(D_all.permute(1, 0, 2).reshape(k, world * fs * cw)[:, :f * cw])
                if checks:
                    ops.max_abs_diff(rview(D), rview(Dn), cplx, result, scratch)
                    if float(result[1].item()) < tol:                                                  # :161, :224
                        return it, Dn, x
                D, Dn = Dn, D
                count += 1
        except KeyboardInterrupt:
            return it, D, x
    return maxiter, D, x


def _allreduce(t, group):
    """Sum over the ranks, in place, also for row-padded (non-contiguous) buffers."""
    return comm.all_reduce_sum(t, group)
