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

from . import comm, ops
from ._device import array_kind, empty2d, full2d, is_torch, np_dtype, require_cuda, to_device2d, to_host
from .utils import assertion

BATCH_METHODS = ['mu']
MINIBATCH_METHODS = ['asg-mu', 'gsg-mu', 'asag-mu', 'gsag-mu', 'svrmu', 'svrmu-acc']
POLL_EVERY = 25
USE_B2B = True     # masked 'l2' x update: ((x D) * M) D^T fused into one kernel where it covers k


def solve(y, D, x=None, tol=1.0e-3, minibatch=None, maxiter=1000, method='mu', likelihood='l2', mask=None,
          random_seed=None, group=None, host_block_rows=None, precision='fp64', **kwargs):
    """NMF, see the module docstring. ``likelihood``: 'l2' | 'gaussian' | 'kl' | 'poisson'.

    ``host_block_rows`` (not in the reference; numpy inputs, 'l2' only): out-of-core mode for data larger than the
    GPU memory.  ``y`` (and ``mask``) stay in host memory, are page-locked in place and streamed through three
    rotating device buffers of that many rows by a copy stream while the previous block is being processed
    (the role of the reference's ``AsyncMinibatchData``, utils/data.py:212-313); ``x`` and ``D`` live on the device.

    ``precision`` (not in the reference): 'fp64' (default; matches the numpy path to ~1e-13) or 'tf32x3' -- full-batch
    'l2' MU, with or without ``mask``, with the large contractions (y D^T, x (D D^T), x^T y / x^T x; masked:
    (x D) * mask, f D^T, x^T (y * mask), x^T f) on the tcgen05 tensor cores, operands split into two TF32 pieces, FP32
    accumulation in tensor memory (sample-axis sums in slabs of 4096 rows added up in FP64); the ratios, the D update
    and the normalisation stay FP64.  Agrees with the FP64 path to ~1e-5 relative on D and x
    (tests/test_tf32x3_gpu.py states the tolerance).  Needs k a multiple of 32 up to 256.

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
    if precision not in ('fp64', 'tf32x3'):
        raise ValueError("precision must be 'fp64' or 'tf32x3', given " + str(precision))
    if precision == 'tf32x3' and (kl or minibatch is not None or host_block_rows is not None
                                  or D.shape[0] % 32 != 0 or D.shape[0] > 256):
        raise NotImplementedError("precision='tf32x3' covers the full-batch 'l2' update with k a multiple of "
                                  "32 up to 256; use precision='fp64'")
    if host_block_rows is not None:
        if minibatch is not None or method != 'mu' or kl or group is not None or is_torch(y) or kwargs:
            raise NotImplementedError("host_block_rows streams numpy data through the full-batch 'mu' / 'l2' solver "
                                      'on one GPU')
        device = require_cuda()
        it, Dd, xd = mu_streamed(y, to_device2d(D, device, copy=True), to_device2d(x, device, copy=True), float(tol),
                                 int(maxiter), mask, int(host_block_rows))
        return it, to_host(Dd, y, np_dtype(y)), to_host(xd, y, np_dtype(y))
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
    it, Dd, xd = mu_device(yd, Dd, xd, float(tol), int(maxiter), kl, md, group=group, precision=precision)
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


USE_SMALL = True
SMALL_MAX_WORK = 5.0e7      # n k f up to which a whole run goes into one cooperative launch (decomp_nmf_mu_small_f64)
GRAPH_MAX_WORK = 2.0e9      # n k f below which a sweep is launch-bound and is replayed from a CUDA graph
GRAPH_MIN_SWEEPS = 12


def mu_device(y, D0, X, tol, maxiter, kl=False, mask=None, group=None, precision='fp64'):
    """Full-batch MU on device tensors; ``X`` [n, k] is updated in place. Returns ``(it, D, X)``.

    Small problems (BASELINE configs[0]: 1000 x 200, k = 20) are bound by the ~12 kernel launches of a sweep, not by
    the kernels: after two ordinary sweeps (which also warm every kernel up) a pair of sweeps (the dictionary
    ping-pongs between two buffers) is captured into a CUDA graph once and replayed.  Nothing in a sweep touches
    the host -- convergence is a device latch whose value is the device-side sweep count under replay -- so the
    results and the returned iteration count are those of the sweep-by-sweep loop."""
    n, f = y.shape
    sweeps = maxiter - 1
    k = D0.shape[0]
    if (USE_SMALL and group is None and not kl and precision == 'fp64' and sweeps >= 1
            and float(n) * f * k <= SMALL_MAX_WORK and ops.nmf_mu_small_supported(n, f, k, mask is not None)):
        # the whole run in one cooperative launch: rows and dictionary stay in shared memory for all sweeps
        dev = y.device
        Dn = empty2d(k, f, False, dev)
        ops.normalize_rows(D0, Dn, False, True)                           # nmf.py:70
        D_out = empty2d(k, f, False, dev)
        it_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = ops.nmf_mu_small(y, mask, X, Dn, D_out, sweeps, tol, it_dev)
        fired = int(it_dev.item()) if tol > 0.0 else 0
        del ws
        return (fired if fired else maxiter), D_out, X
    solver = MuSolver(y, D0, X, tol, kl=kl, mask=mask, group=group, precision=precision)
    stopped_at = 0
    it = 1
    if (group is None and not solver.tf32 and sweeps >= GRAPH_MIN_SWEEPS
            and float(n) * f * D0.shape[0] <= GRAPH_MAX_WORK):
        solver.counted = True
        for it in (1, 2):
            solver.sweep(it)
        it = 3
        cur = torch.cuda.current_stream(y.device)
        side = torch.cuda.Stream(device=y.device)
        side.wait_stream(cur)
        graph = torch.cuda.CUDAGraph()
        before = ops.LAUNCHES
        # (capture_begin / capture_end directly: the torch.cuda.graph context manager also runs the garbage collector
        # and empties the allocator cache, which costs more than the sweeps it is supposed to save)
        with torch.cuda.stream(side):
            graph.capture_begin()
            try:
                solver.sweep(3)
                solver.sweep(4)
            finally:
                graph.capture_end()
        per_replay = ops.LAUNCHES - before
        cur.wait_stream(side)
        since_poll = 2
        while it + 1 <= sweeps:
            graph.replay()
            ops._count(per_replay)
            it += 2
            since_poll += 2
            if solver.checks and since_poll >= POLL_EVERY:
                since_poll = 0
                stopped_at = solver.fired()
                if stopped_at:
                    break
    while not stopped_at and it <= sweeps:
        if solver.checks and it % POLL_EVERY == 0:
            stopped_at = solver.fired()
            if stopped_at:
                break
        solver.sweep(it)
        it += 1
    if solver.checks and not stopped_at:
        stopped_at = solver.fired()
    if stopped_at:
        return stopped_at, solver.Dbuf[stopped_at % 2], X
    return maxiter, solver.Dbuf[max(maxiter - 1, 0) % 2], X


class MuSolver(object):
    """Device state of a full-batch MU run; ``sweep(it)`` enqueues sweep number ``it`` (1-based, reading
    ``Dbuf[(it - 1) % 2]`` and writing ``Dbuf[it % 2]``) on the current stream without synchronising."""

    def __init__(self, y, D0, X, tol, kl=False, mask=None, group=None, precision='fp64'):
        dev = y.device
        self.y, self.X, self.mask, self.kl, self.tol, self.group = y, X, mask, kl, tol, group
        self.tf32 = precision == 'tf32x3'
        if self.tf32 and (kl or D0.shape[0] % 32 != 0 or D0.shape[0] > 256):
            raise NotImplementedError("precision='tf32x3': 'l2' update, k a multiple of 32 up to 256")
        self.n, self.f = n, f = y.shape
        self.k = k = D0.shape[0]
        self.Dbuf = [empty2d(k, f, False, dev), empty2d(k, f, False, dev)]
        ops.normalize_rows(D0, self.Dbuf[0], False, True)                 # nmf.py:70
        self.Draw = empty2d(k, f, False, dev)
        self.Dt = empty2d(f, k, False, dev)
        self.POS = empty2d(k, f, False, dev)
        if not self.tf32:
            self.NEG = empty2d(n, k, False, dev)
            self.ws = ops.gemm_tn_workspace_for([(k, f, n), (k, k, n)], dev)
        else:
            # TF32 pairs (float32 hi + lo) of everything the tensor cores read: y and y^T once, x / x^T / D / D D^T
            # per sweep.  All operands are K-major: x^T y is the NT product of x^T [k, n] and y^T [f, n].
            # x^T and y^T are stored K-blocked ([n / 4096][rows][4096]) so that a slab of the sample-axis contraction
            # is a compact piece of memory instead of one 128-byte line per row, 4 n bytes apart.
            blk = ops.TF32_K_PER_SPLIT
            ysrc = y
            if mask is not None:
                # masked model (grads.py:112-115, 122-125): y * mask replaces y in both numerators, once; the mask
                # itself is read by the F = (x D) * mask epilogue as FP32
                ysrc = empty2d(n, f, False, dev)
                ops.mask_mul(y, mask, ysrc)
                self.M32 = ops.to_f32(mask)
                self.Fh, self.Fl = ops.empty_f32(n, f, dev), ops.empty_f32(n, f, dev)
                self.FTh = ops.empty_f32_blocked(n, f, blk, dev, zero_tail=True)
                self.FTl = ops.empty_f32_blocked(n, f, blk, dev, zero_tail=True)
                self.DTh, self.DTl = ops.empty_f32(f, k, dev), ops.empty_f32(f, k, dev)
                self.NEGD = empty2d(k, f, False, dev)
            self.Yh, self.Yl = ops.split_tf32(ysrc)
            self.YTh, self.YTl = ops.split_transpose_tf32(ysrc, block=blk)
            del ysrc
            self.Xh, self.Xl = ops.split_tf32(X)
            self.XTh = ops.empty_f32_blocked(n, k, blk, dev, zero_tail=True)
            self.XTl = ops.empty_f32_blocked(n, k, blk, dev, zero_tail=True)
            self.Dh, self.Dl = ops.empty_f32(k, f, dev), ops.empty_f32(k, f, dev)
            self.Gh, self.Gl = ops.empty_f32(k, k, dev), ops.empty_f32(k, k, dev)
            self.NEG32 = ops.empty_f32(n, k, dev)
            ws_bytes = max(ops._lib.lib().decomp_gemm_nt_tf32x3_splitk_workspace_bytes(k, max(f, k), n, blk),
                           ops._lib.lib().decomp_gemm_nt_tf32x3_splitk_workspace_bytes(k, k, f, 512))
            self.ws32 = ops.workspace(ws_bytes, dev)
            self.ws = None
        self.checks = tol > 0.0
        self.comm_events = None   # set to [] to collect (start, end) CUDA event pairs around each sweep's all-reduces
        self.latch = torch.zeros(1, dtype=torch.int32, device=dev) if self.checks else None
        self.scratch = torch.zeros(2, dtype=torch.int32, device=dev)   # ticket counter, sweep counter
        self.counted = False    # True: the latch value is the device-side sweep count (graph replay)
        self.maxdiff = torch.zeros(2, dtype=torch.float64, device=dev)
        self.ym = y
        if mask is not None and not self.tf32:
            self.ym = empty2d(n, f, False, dev)
            ops.mask_mul(y, mask, self.ym)                                # y * mask, once (grads.py:113,123)
        self.b2b = bool(USE_B2B and mask is not None and not kl and not self.tf32 and ops.gemm_b2b_masked_supported(k))
        if self.tf32:
            if mask is None:
                self.G = empty2d(k, k, False, dev)
                self.S = empty2d(k, k, False, dev)
        elif mask is not None or kl:
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

    def _comm_mark(self, start=None):
        """Timing hook of bench.py: CUDA events around the all-reduces of a sweep (only when comm_events is a list)."""
        if self.comm_events is None:
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        if start is not None:
            self.comm_events.append([(start, ev)])
        return ev

    def sweep(self, it):
        E = ops.epilogue
        y, ym, X, mask, latch, ws, group = self.y, self.ym, self.X, self.mask, self.latch, self.ws, self.group
        D, Dn = self.Dbuf[(it - 1) % 2], self.Dbuf[it % 2]
        Dt, POS, Draw = self.Dt, self.POS, self.Draw
        NEG = None if self.tf32 else self.NEG
        if self.tf32 and mask is not None:
            # ---- masked l2 (grads.py:112-115, 122-125) on the tcgen05 tensor cores; the [n, f] intermediate
            # f = (x D) * mask travels as a TF32 pair: row-major for f D^T, K-blocked transposed for x^T f
            ops.split_tf32(D, self.Dh, self.Dl)
            ops.split_transpose_tf32(D, self.DTh, self.DTl)
            ops.gemm_nt_mask_tf32x3(self.Xh, self.Xl, self.DTh, self.DTl, self.M32, F=(self.Fh, self.Fl), skip=latch)
            ops.gemm_nt_tf32x3(self.Fh, self.Fl, self.Dh, self.Dl, self.NEG32, skip=latch)
            ops.nmf_xupdate_tf32x3(self.Yh, self.Yl, self.Dh, self.Dl, X, self.NEG32, self.Xh, self.Xl, self.XTh,
                                   self.XTl, skip=latch)
            ops.gemm_nt_mask_tf32x3(self.Xh, self.Xl, self.DTh, self.DTl, self.M32, FT=(self.FTh, self.FTl), skip=latch)
            ops.gemm_nt_tf32x3_splitk(self.XTh, self.XTl, self.YTh, self.YTl, POS, self.ws32, skip=latch, K=self.n)
            ops.gemm_nt_tf32x3_splitk(self.XTh, self.XTl, self.FTh, self.FTl, self.NEGD, self.ws32, skip=latch, K=self.n)
            if group is not None:
                t0 = self._comm_mark()
                _allreduce2d(POS, group)
                _allreduce2d(self.NEGD, group)
                self._comm_mark(t0)
            ops.mu_update(D, POS, self.NEGD, Draw, skip=latch)
        elif self.tf32:
            G, S = self.G, self.S
            # ---- x update (grads.py:108-111, f.dot(d.T) re-associated) on the tcgen05 tensor cores
            ops.split_tf32(D, self.Dh, self.Dl)
            ops.gemm_nt_tf32x3_splitk(self.Dh, self.Dl, self.Dh, self.Dl, G, self.ws32, k_per_split=512, skip=latch)
            ops.split_tf32(G, self.Gh, self.Gl)
            ops.gemm_nt_tf32x3(self.Xh, self.Xl, self.Gh, self.Gl, self.NEG32, skip=latch)
            ops.nmf_xupdate_tf32x3(self.Yh, self.Yl, self.Dh, self.Dl, X, self.NEG32, self.Xh, self.Xl, self.XTh,
                                   self.XTl, skip=latch)
            # ---- D update (grads.py:117-121): statistics over the sample axis, FP32 slabs of 4096 rows summed in FP64
            ops.gemm_nt_tf32x3_splitk(self.XTh, self.XTl, self.YTh, self.YTl, POS, self.ws32, skip=latch, K=self.n)
            ops.gemm_nt_tf32x3_splitk(self.XTh, self.XTl, self.XTh, self.XTl, S, self.ws32, skip=latch, K=self.n)
            if group is not None:
                t0 = self._comm_mark()
                _allreduce2d(POS, group)
                _allreduce2d(S, group)
                self._comm_mark(t0)
            ops.make_rhs(D, False, False, out=Dt, skip=latch)
            ops.gemm_nt(S, Dt, E(ops.EPI_MU_DEN, Draw, x=D, other=POS), skip=latch)
        elif not self.kl and mask is None:
            G, S = self.G, self.S
            # ---- x update (grads.py:108-111 with f.dot(d.T) re-associated)
            ops.gemm_nt(D, D, E(ops.EPI_STORE, G), skip=latch)
            ops.gemm_nt(X, G, E(ops.EPI_STORE, NEG), skip=latch)
            ops.gemm_nt(y, D, E(ops.EPI_MU_NUM, X, x=X, other=NEG), skip=latch)
            # ---- D update (grads.py:117-121): sufficient statistics over the sample axis
            ops.gemm_tn(X, y, POS, workspace=ws, skip=latch)
            ops.gemm_tn(X, X, S, workspace=ws, skip=latch)
            if group is not None:
                t0 = self._comm_mark()
                _allreduce2d(POS, group)
                _allreduce2d(S, group)
                self._comm_mark(t0)
            ops.make_rhs(D, False, False, out=Dt, skip=latch)
            ops.gemm_nt(S, Dt, E(ops.EPI_MU_DEN, Draw, x=D, other=POS), skip=latch)
        else:
            F, NEGD = self.F, self.NEGD
            ops.make_rhs(D, False, False, out=Dt, skip=latch)
            if not self.kl:
                # ---- masked l2 (grads.py:112-115, 122-125)
                if self.b2b:
                    ops.gemm_b2b_masked(X, Dt, E(ops.EPI_STORE, NEG, mask=mask), skip=latch)    # ((x D) * M) D^T fused
                else:
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
                        comm.all_reduce_sum(self.xsum, group)
                    ops.scale(self.ones_kf, NEGD, rowscale=self.xsum)
                else:
                    ops.gemm_tn(X, mask, NEGD, workspace=ws, skip=latch)
            if group is not None:
                t0 = self._comm_mark()
                _allreduce2d(POS, group)
                if not (self.kl and mask is None):
                    _allreduce2d(NEGD, group)
                self._comm_mark(t0)
            ops.mu_update(D, POS, NEGD, Draw, skip=latch)
        # ---- l2_strict + max|D - D_new| < tol (batch_mu.py:21-23)
        ops.normalize_rows(Draw, Dn, False, True, D_ref=D if self.checks else None, tol=self.tol, latch=latch,
                           latch_value=-1 if self.counted else it, maxdiff=self.maxdiff if self.checks else None,
                           scratch=self.scratch, skip=latch)


class _PinnedRows(object):
    """Host rows page-locked in place (cudaHostRegister) for asynchronous block copies; released on close()."""

    def __init__(self, a):
        # float32 rows are staged as they are and widened by the device-side copy: no float64 host copy of the data
        a = np.ascontiguousarray(a)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        self.t = torch.from_numpy(a)
        self.registered = False
        if not self.t.is_pinned():
            rc = torch.cuda.cudart().cudaHostRegister(self.t.data_ptr(), self.t.numel() * self.t.element_size(), 0)
            self.registered = int(rc) == 0
        self.cols = a.shape[1]

    def close(self):
        if self.registered:
            torch.cuda.cudart().cudaHostUnregister(self.t.data_ptr())
            self.registered = False


def mu_streamed(y, D0, X, tol, maxiter, mask, block_rows):
    """Full-batch 'l2' MU with ``y`` / ``mask`` streamed from host memory in row blocks; ``X`` [n, k] and the
    dictionary stay on the device.  One H2D pass over ``y`` per sweep: the x update and the statistics of a
    block are both computed while it is resident.  Returns ``(it, D, X)``."""
    dev = X.device
    n, f = y.shape
    k = D0.shape[0]
    masked = mask is not None
    rows = max(1, min(block_rows, n))
    nblk = (n + rows - 1) // rows
    hy = _PinnedRows(y)
    hm = _PinnedRows(mask) if masked else None
    NB = 3
    ybuf = [empty2d(rows, f, False, dev) for _ in range(NB)]
    mbuf = [empty2d(rows, f, False, dev) for _ in range(NB)] if masked else None
    ready = [torch.cuda.Event() for _ in range(NB)]      # block has arrived in ybuf[i]
    freed = [torch.cuda.Event() for _ in range(NB)]      # compute has finished with ybuf[i]
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    # the staging buffers come from the caching allocator of the main stream: earlier users of those blocks that are
    # still queued there must have finished before the copy stream writes into them
    copy_stream.wait_stream(main)
    E = ops.epilogue

    Dbuf = [empty2d(k, f, False, dev), empty2d(k, f, False, dev)]
    ops.normalize_rows(D0, Dbuf[0], False, True)
    Draw, Dt, POS = empty2d(k, f, False, dev), empty2d(f, k, False, dev), empty2d(k, f, False, dev)
    NEG = empty2d(rows, k, False, dev)
    ws = ops.gemm_tn_workspace_for([(k, f, rows), (k, k, rows)], dev)
    if masked:
        F, YM, NEGD = empty2d(rows, f, False, dev), empty2d(rows, f, False, dev), empty2d(k, f, False, dev)
    else:
        G, S = empty2d(k, k, False, dev), empty2d(k, k, False, dev)
    result = torch.zeros(2, dtype=torch.float64, device=dev)
    scratch = torch.zeros(1, dtype=torch.int32, device=dev)

    def fetch(b, slot, first_use):
        r0, r1 = b * rows, min(n, (b + 1) * rows)
        with torch.cuda.stream(copy_stream):
            if not first_use:
                copy_stream.wait_event(freed[slot])
            ybuf[slot][:r1 - r0].copy_(hy.t[r0:r1], non_blocking=True)
            if masked:
                mbuf[slot][:r1 - r0].copy_(hm.t[r0:r1], non_blocking=True)
            ready[slot].record(copy_stream)

    try:
        issued = 0
        for it in range(1, maxiter):
            D, Dn = Dbuf[(it - 1) % 2], Dbuf[it % 2]
            ops.make_rhs(D, False, False, out=Dt)
            if not masked:
                ops.gemm_nt(D, D, E(ops.EPI_STORE, G))
            for b in range(min(NB - 1, nblk)):                       # prefetch
                fetch(b, issued % NB, issued < NB)
                issued += 1
            base = issued - min(NB - 1, nblk)
            for b in range(nblk):
                if b + NB - 1 < nblk:
                    fetch(b + NB - 1, issued % NB, issued < NB)
                    issued += 1
                slot = (base + b) % NB
                r0, r1 = b * rows, min(n, (b + 1) * rows)
                r = r1 - r0
                main.wait_event(ready[slot])
                yb, xb = ybuf[slot][:r], X[r0:r1]
                comb, beta = (0, 0.0) if b == 0 else (1, 1.0)        # accumulate the statistics over the blocks
                if not masked:
                    ops.gemm_nt(xb, G, E(ops.EPI_STORE, NEG[:r]))
                    ops.gemm_nt(yb, D, E(ops.EPI_MU_NUM, xb, x=xb, other=NEG[:r]))
                    ops.gemm_tn(xb, yb, POS, combine=comb, beta=beta, workspace=ws)
                    ops.gemm_tn(xb, xb, S, combine=comb, beta=beta, workspace=ws)
                else:
                    mb = mbuf[slot][:r]
                    ops.mask_mul(yb, mb, YM[:r])
                    ops.gemm_nt(xb, Dt, E(ops.EPI_STORE_MASK, F[:r], mask=mb))
                    ops.gemm_nt(F[:r], D, E(ops.EPI_STORE, NEG[:r]))
                    ops.gemm_nt(YM[:r], D, E(ops.EPI_MU_NUM, xb, x=xb, other=NEG[:r]))
                    ops.gemm_nt(xb, Dt, E(ops.EPI_STORE_MASK, F[:r], mask=mb))
                    ops.gemm_tn(xb, YM[:r], POS, combine=comb, beta=beta, workspace=ws)
                    ops.gemm_tn(xb, F[:r], NEGD, combine=comb, beta=beta, workspace=ws)
                freed[slot].record(main)
            if not masked:
                ops.gemm_nt(S, Dt, E(ops.EPI_MU_DEN, Draw, x=D, other=POS))
            else:
                ops.mu_update(D, POS, NEGD, Draw)
            ops.normalize_rows(Draw, Dn, False, True)
            if tol > 0.0:
                ops.max_abs_diff(D, Dn, False, result, scratch)
                if float(result[1].item()) < tol:
                    return it, Dn, X
        torch.cuda.synchronize(dev)
        return maxiter, Dbuf[max(maxiter - 1, 0) % 2], X
    finally:
        torch.cuda.synchronize(dev)
        hy.close()
        if hm is not None:
            hm.close()


def _allreduce2d(t, group):
    """Sum a (possibly row-padded) 2-D statistic over the ranks, in place."""
    return comm.all_reduce_sum(t, group)
