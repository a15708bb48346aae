"""Batched Lasso by ISTA / FISTA / accelerated ISTA on the B200.

Drop-in for the reference's ``decomp.lasso`` (``solve`` :19-94, ``solve_fastpath`` :97-189):
same signatures, defaults, validation errors and return tuple ``(it, x)``.

    argmin_x  1/(2n) |y - x A|^2 + alpha |x|      y [..., f],  x [..., k],  A [k, f]

Device data flow (all FP64; complex data as interleaved doubles through the 2x2 real embedding):

    prologue   s_k = |A_k|, A <- A/s, alpha_k, tol_k, x <- x*s               lasso.py:120-138,163
               G = A A^H (mask: (A * mean(mask)) A^H), 1/L by Gershgorin       lasso.py:285-287,317-319
               yAh = y A^H (mask: (y*mask) A^H)                                lasso.py:289,321
    iteration  x_new = shrink(w + (yAh - w G)/L, alpha/L)  [+ momentum, + convergence latch]     lasso.py:244-256,405-414
               no per-problem mask, width <= 256 doubles: the iterate stays on chip and one launch of the resident
               kernel runs up to 32 iterations (ops.lasso_resident); otherwise one NT GEMM launch per iteration whose
               epilogue does the whole update (per-problem mask: w A -> *mask fused in a first GEMM's epilogue)
    epilogue   x / s                                                           lasso.py:189

The convergence test (every 10th iteration, global over the batch, lasso.py:293/409) is evaluated inside
the update launch (launches are cut so that it falls on their last iteration); when it passes, a device latch is
set and every later launch returns immediately, so the host never has to synchronise inside the loop to get the
reference's result.  Host arrays with ``tol <= 0`` are solved in row chunks whose copies overlap the iterations.
"""
import math
import queue
import threading

import numpy as np
import torch

from . import comm, ops
from ._device import (array_kind, empty2d, flatten_rows, is_torch, np_dtype, require_cuda, to_device1d,
                      to_device2d, to_host)
from ._lib import rview
from .utils import assertion

AVAILABLE_METHODS = ['ista', 'cd', 'acc_ista', 'fista', 'parallel_cd', 'admm']
AVAILABLE_NNLS_METHODS = ['ista_pos', 'cd_pos', 'acc_ista_pos', 'fista_pos', 'parallel_cd_pos', 'admm_pos']
DEVICE_RULES = ('ista', 'fista', 'acc_ista')
RESIDENT_PAD_WORK = 1.5e8   # rows x width^2 below which one iteration is launch-bound (< ~10 us of DMMA)
RESIDENT_MIN_ITERS = 7      # launches up to this length run per iteration when the batch is not launch-bound
USE_B2B = True        # masked iteration: ((w A) * M) A^H fused into one kernel where it covers the width
USE_RESIDENT = True   # several iterations per launch with the iterate on chip where the kernel covers the shape
POLL_EVERY = 50   # iterations between (cheap) host reads of the convergence latch


def solve(y, A, alpha, x=None, tol=1.0e-3, method='ista', maxiter=1000, mask=None, precision='fp64', **kwargs):
    """Solve the batched Lasso problem; see the module docstring. Returns ``(it, x)``.

    Arguments are numpy arrays (result: numpy) or CUDA torch tensors (result: torch, no host copy).
    ``method``: 'ista' | 'fista' | 'acc_ista', optionally suffixed '_pos' for non-negative x.
    The reference's sequential / inverse-based methods ('cd', 'parallel_cd', 'admm') are outside the
    hot path and raise ``NotImplementedError``.

    ``precision`` (not in the reference): 'fp64' (default; matches the numpy path to ~1e-13) or 'tf32x3' -- the
    iteration's GEMMs on the tcgen05 tensor cores with operands split into two TF32 pieces and FP32 accumulation
    (the update itself stays FP64); agrees with the FP64 path to ~1e-5 relative on x.  Without a per-problem mask:
    w (I - G/L), k (2k for complex data) a multiple of 32 up to 256; with one: (w A) * mask and its product with A^H,
    the [B, f] intermediate travelling as a TF32 pair, any even real width.
    """
    array_kind(y, A, x, mask)
    # x = None means zeros(y.shape[:-1] + (k,), y.dtype) (lasso.py:73-74); they are created on the device

    assertion.assert_dtypes(y=y, A=A, x=x)
    assertion.assert_dtypes(mask=mask, dtypes='f')
    assertion.assert_nonnegative(mask)
    assertion.assert_ndim('A', A, ndim=2)
    assertion.assert_shapes('x', x, 'A', A, axes=1)
    if x is not None:
        assertion.assert_shapes('y', y, 'x', x, axes=list(range(x.ndim - 1)))
    assertion.assert_shapes('y', y, 'A', A, axes=[-1])
    if mask is not None and mask.ndim == 1:
        assertion.assert_shapes('y', y, 'mask', mask, axes=[-1])
    else:
        assertion.assert_shapes('y', y, 'mask', mask)
    if method not in AVAILABLE_METHODS + AVAILABLE_NNLS_METHODS:
        raise ValueError('Available methods are {0:s}. Given {1:s}'.format(str(AVAILABLE_METHODS), method))
    assert np_dtype(A).kind != 'c' or method[-4:] != '_pos'
    return solve_fastpath(y, A, alpha, x, tol, maxiter, method, None, mask=mask, precision=precision, **kwargs)


def solve_fastpath(y, A, alpha, x, tol, maxiter, method, xp=None, mask=None, group=None, precision='fp64',
                   **kwargs):
    """Assertion-free entry (reference: decomp/lasso.py:97-189). ``xp`` is accepted and ignored.

    ``group``: optional ``torch.distributed`` process group; the batch rows given to this rank are its
    shard, ``A`` is replicated, and only the convergence decision is exchanged (one int32 every 10th
    iteration), so that every rank stops at the same iteration as the single-device run would.
    """
    positive = method[-4:] == '_pos'
    rule = method[:-4] if positive else method
    if rule not in DEVICE_RULES:
        if rule in AVAILABLE_METHODS:
            raise NotImplementedError('Method ' + method + ' is not on the B200 hot path '
                                      '(ista, fista, acc_ista and their _pos variants are).')
        raise NotImplementedError('Method ' + method + ' is not yet implemented.')
    if kwargs:
        raise TypeError('unexpected keyword arguments ' + str(sorted(kwargs)))
    if np.ndim(alpha) != 0:
        raise NotImplementedError('alpha must be a scalar')

    device = require_cuda()
    out_dtype = np_dtype(y)
    batch_shape = tuple(y.shape[:-1])
    k = A.shape[0]
    # (a per-problem mask couples the rows through the batch mean of the mask, lasso.py:300-303: one piece)
    if group is None and not float(tol) > 0.0 and not is_torch(y) and (mask is None or mask.ndim == 1):
        chunks = _row_chunks(flatten_rows(y).shape[0], y.shape[-1], k * (2 if np_dtype(A).kind == 'c' else 1), device,
                             pinned=_is_pinned(y))
        if chunks is not None:
            it, res = _solve_pipelined(flatten_rows(y), A, float(alpha), None if x is None else flatten_rows(x),
                                       int(maxiter), rule, positive, mask, precision, chunks, device, out_dtype)
            return it, res.reshape(batch_shape + (k,))
    y2 = to_device2d(flatten_rows(y), device, copy=False)
    x2 = to_device2d(flatten_rows(x), device, copy=False) if x is not None else None
    A2 = to_device2d(A, device, copy=False)
    m2 = None
    if mask is not None:
        m2 = to_device1d(mask, device) if mask.ndim == 1 else to_device2d(flatten_rows(mask), device, copy=False)

    state = lasso_device(y2, A2, float(alpha), x2, float(tol), int(maxiter), rule, positive, m2, group=group,
                         precision=precision)
    it = state.iterations()
    res = to_host(state.result, y, out_dtype)
    return it, res.reshape(batch_shape + (k,))


PIPELINE_MIN_BYTES = 64 << 20   # host batches below this are solved in one piece
PIPELINE_MAX_CHUNK_BYTES = 2 << 30   # upper bound of one chunk of y on the device
PIPELINE_DEPTH = 3              # chunks resident on the device at a time: uploading / iterating / downloading


def _row_chunks(B, f, k_cols, device, pinned=True):
    """Row ranges for the pipelined host path, or None when the batch is too small to be worth splitting.

    Chunks are whole rounds of the persistent GEMM grid (one CTA tile of 128 rows x 64 columns per SM and round), so
    splitting costs no tile quantisation; a short first chunk lets the iterations start early and a short last one
    keeps the final device-to-host copy small (the ragged rest of the batch rides in a long middle chunk).

    Uploads from pageable host memory keep host threads busy (memcpy into page-locked staging buffers) and take about
    as long as the iterations they are supposed to hide behind: such a batch is cut into a short first chunk and
    uniform chunks of four rounds, uploaded by a helper thread while the calling thread enqueues the iterations."""
    if B * f * 8 < PIPELINE_MIN_BYTES:
        return None
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    tiles_n = -(-k_cols // 64)
    unit = 128 * max(1, sms // tiles_n)
    units = B // unit
    if units < 4:
        return None
    if not pinned:
        # a short first chunk (its upload is the only one nothing hides), then chunks of four rounds: the uploads run in
        # a helper thread, so the chunk count only costs enqueuing work
        out, r0, step = [], 0, 2 * unit
        while r0 < B:
            out.append((r0, min(B, r0 + step)))
            r0 += step
            step = 4 * unit
        if len(out) > 1 and out[-1][1] - out[-1][0] < unit:      # a short tail joins its neighbour
            out[-2:] = [(out[-2][0], B)]
        return out
    first = 2 * unit if units >= 8 else unit
    middle = B - first - unit                  # whole rounds plus the ragged rest of the batch
    pieces = max(2 if middle >= 8 * unit else 1, -(-middle * f * 8 // PIPELINE_MAX_CHUNK_BYTES))
    pieces = max(1, min(pieces, middle // unit))
    head = (middle // unit // pieces) * unit
    sizes = [first] + [head] * (pieces - 1) + [middle - head * (pieces - 1), unit]
    out, r0 = [], 0
    for n in sizes:
        out.append((r0, r0 + n))
        r0 += n
    assert r0 == B
    return out


def _is_pinned(a):
    """True if the numpy array lives in page-locked host memory (its copies are asynchronous and at PCIe speed)."""
    try:
        return bool(a.size > 0 and torch.from_numpy(a[:1]).is_pinned())      # a view: nothing is copied
    except (TypeError, ValueError, RuntimeError):
        return False


_COPY_STREAMS = {}


def _copy_streams(device):
    """One upload and one download stream per device, created on first use."""
    key = (device.type, device.index)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = (torch.cuda.Stream(device), torch.cuda.Stream(device))
    return _COPY_STREAMS[key]


def _solve_pipelined(y, A, alpha, x, maxiter, rule, positive, mask, precision, chunks, device, out_dtype):
    """Host arrays, ``tol <= 0``: every row runs exactly ``maxiter - 1`` iterations whatever the other rows do
    (lasso.py:293/409 never fires), so the batch is solved chunk by chunk with the upload of the next chunk and the
    download of the previous one overlapping the iterations of the current one.  Row results are bitwise those of the
    one-piece solve: a row's dot products do not depend on which tile it sits in.  Only PIPELINE_DEPTH chunks of at
    most PIPELINE_MAX_CHUNK_BYTES are on the device at a time, so the host batch may exceed the device memory.

    Uploads from ordinary (pageable) memory keep the host busy (memcpy into page-locked staging buffers, see
    _device._staged_upload), so they run in a helper thread that works ahead of the thread enqueuing the iterations;
    page-locked inputs are copied asynchronously from the enqueuing thread itself."""
    cur = torch.cuda.current_stream(device)
    up, down = _copy_streams(device)
    up.wait_stream(cur)
    down.wait_stream(cur)
    k = A.shape[0]
    A2 = to_device2d(A, device, copy=False)
    m1 = to_device1d(mask, device) if mask is not None else None
    tdt = getattr(torch, np.dtype(out_dtype).name)
    host = torch.empty((y.shape[0], k), dtype=tdt, pin_memory=True)
    # At most PIPELINE_DEPTH chunks live on the device (uploading / iterating / downloading), so the batch may be
    # larger than HBM: chunk j is uploaded once chunk j - PIPELINE_DEPTH has been downloaded and its buffers dropped.
    n = len(chunks)
    finished = {}

    def upload(j):
        r0, r1 = chunks[j]
        with torch.cuda.stream(up):
            yc = to_device2d(y[r0:r1], device, copy=False)
            xc = to_device2d(x[r0:r1], device, copy=False) if x is not None else None
            ev = torch.cuda.Event()
            ev.record(up)
        return yc, xc, ev

    def retire(j):
        """Blocks until chunk j has been downloaded, then drops its device buffers."""
        if j >= 0:
            finished.pop(j)[0].synchronize()

    threaded = not _is_pinned(y)
    if threaded:
        ready, issued = queue.Queue(), [threading.Event() for _ in chunks]
        abort = threading.Event()

        def uploader():
            try:
                torch.cuda.set_device(device)
                for j in range(n):
                    if j - PIPELINE_DEPTH >= 0:
                        while not issued[j - PIPELINE_DEPTH].wait(0.05):
                            if abort.is_set():
                                return
                        retire(j - PIPELINE_DEPTH)
                    if abort.is_set():
                        return
                    ready.put(upload(j))
            except BaseException as exc:           # surfaces in the enqueuing thread
                ready.put(exc)

        worker = threading.Thread(target=uploader, name='decomp-upload', daemon=True)
        worker.start()
    else:
        staged = {0: upload(0)}

    try:
        for c, (r0, r1) in enumerate(chunks):
            if threaded:
                item = ready.get()
                if isinstance(item, BaseException):
                    raise item
                yc, xc, ev = item
            else:
                yc, xc, ev = staged.pop(c)
            cur.wait_event(ev)
            state = lasso_device(yc, A2, alpha, xc, 0.0, maxiter, rule, positive, m1, precision=precision,
                                 rows_hint=y.shape[0])      # same kernel choice as the one-piece solve
            res = state.result.to(dtype=tdt)
            done = torch.cuda.Event()
            done.record(cur)
            down.wait_event(done)
            with torch.cuda.stream(down):
                host[r0:r1].copy_(res, non_blocking=True)
                copied = torch.cuda.Event()
                copied.record(down)
            finished[c] = (copied, yc, xc, state, res)      # the buffers stay referenced until the copy has finished
            del yc, xc, state, res
            if threaded:
                issued[c].set()
            elif c + 1 < n:
                # the next chunk goes up while this one iterates
                retire(c + 1 - PIPELINE_DEPTH)
                staged[c + 1] = upload(c + 1)
    except BaseException:
        if threaded:
            abort.set()
        raise
    if threaded:
        worker.join()
    down.synchronize()
    finished.clear()
    return maxiter - 1, host.numpy()


def _padded2d(rows, cols, cplx, width, device):
    """Zero-filled [rows, cols] buffer (complex if ``cplx``) inside a real [rows, width] one; returns (view, base)."""
    base = torch.zeros((max(rows, 1), width), dtype=torch.float64, device=device)[:rows]
    if cplx:
        return torch.view_as_complex(base.view(rows, width // 2, 2))[:, :cols], base
    return base[:, :cols], base


class LassoState(object):
    """Handle on an enqueued device solve: ``result`` [B, k] and the lazily read iteration count."""

    def __init__(self, result, latch, maxiter):
        self.result = result
        self.latch = latch
        self.maxiter = maxiter

    def iterations(self):
        fired = int(self.latch.item()) if self.latch is not None else 0
        return fired - 1 if fired > 0 else self.maxiter - 1


def _momentum_schedule(rule, maxiter):
    """Extrapolation weight applied after iteration i (lasso.py:411-412 fista, :355 acc_ista)."""
    if rule == 'ista':
        return [0.0] * maxiter
    if rule == 'acc_ista':
        return [i / (i + 3) for i in range(maxiter)]
    out, beta = [], 1.0
    for _ in range(maxiter):
        beta_next = 0.5 * (1.0 + math.sqrt(1.0 + 4.0 * beta * beta))
        out.append((beta - 1.0) / beta_next)
        beta = beta_next
    return out


def lasso_device(y, A, alpha, x, tol, maxiter, rule, positive, mask=None, out=None, group=None, precision='fp64',
                 rows_hint=None):
    """Enqueue a whole solve on the current stream. All arguments are device tensors ([B, f], [k, f],
    [B, k] or None for zeros; mask None, [f] or [B, f]). Nothing is synchronised unless the latch has to
    be polled (``tol > 0`` and more than POLL_EVERY iterations). Returns a ``LassoState``."""
    solver = LassoSolver(y, A, alpha, x, tol, maxiter, rule, positive, mask=mask, group=group, precision=precision,
                         rows_hint=rows_hint)
    solver.iterate(0, solver.n_inplace)
    return solver.finish(out)


class LassoSolver(object):
    """One batched Lasso solve, split into set-up / iterations / read-out so that callers (and bench.py) can
    enqueue exactly the iterations they want. Every method only enqueues work on the current stream."""

    def __init__(self, y, A, alpha, x, tol, maxiter, rule, positive, mask=None, group=None, precision='fp64',
                 rows_hint=None):
        dev = y.device
        if precision not in ('fp64', 'tf32x3'):
            raise ValueError("precision must be 'fp64' or 'tf32x3', given " + str(precision))
        self.tf32 = precision == 'tf32x3'
        self.cplx = cplx = A.is_complex()
        self.cw = cw = 2 if cplx else 1
        self.B, self.f = B, f = y.shape
        self.k = k = A.shape[0]
        self.rule, self.maxiter, self.group, self.mask = rule, maxiter, group, mask
        yr, Ar = rview(y), rview(A)
        self.full_mask = full_mask = mask is not None and mask.dim() == 2
        self.b2b = False
        self.shrink = ops.SHRINK_POSITIVE if positive else (ops.SHRINK_COMPLEX if cplx else ops.SHRINK_REAL)

        # ---- prologue (lasso.py:120-138, 163)
        mult_dev = None
        if mask is not None and not full_mask:
            Am, ym = empty2d(k, f, cplx, dev), empty2d(B, f, cplx, dev)
            ops.scale(Ar, rview(Am), cwidth=cw, colscale=mask)
            ops.scale(yr, rview(ym), cwidth=cw, colscale=mask)
            Ar, yr = rview(Am), rview(ym)
            mult_dev = ops.row_sums(mask.view(1, f))
        self.s = s = ops.row_norms(Ar, cplx)
        An = empty2d(k, f, cplx, dev)
        ops.scale(Ar, rview(An), cwidth=cw, rowscale=s, invert_row=True)
        Anr = rview(An)
        self.alpha_vec, self.tol_vec = ops.lasso_vectors(s, alpha, tol, mult=1.0 if full_mask else float(f),
                                                         mult_dev=mult_dev)
        # iterate resident on chip, several iterations per launch: FP64, no per-problem mask, ista / fista, problem
        # width (in doubles) 32 / 64 / 128 / 256 -- narrower problems are zero-padded up to the next of those when
        # that costs little (<= 1.3x the flops) or when the iteration is launch-bound anyway
        n_real = k * cw
        self.rows_total = rows_hint or B          # rows of the whole batch when this solver runs one chunk of it
        if group is not None and rows_hint is None:
            # the kernel choice below must not depend on the size of this rank's shard: use the largest shard
            rows = torch.tensor([B], dtype=torch.int64, device=dev)
            torch.distributed.all_reduce(rows, op=torch.distributed.ReduceOp.MAX, group=group)
            self.rows_total = int(rows.item())
        self.npad = next((w for w in (32, 64, 128, 256) if w >= n_real), 0)
        self.resident = bool(USE_RESIDENT and not self.tf32 and not full_mask and rule in ('ista', 'fista', 'acc_ista')
                             and self.npad and ops.lasso_resident_supported(self.npad)
                             and (self.npad == n_real or (self.npad / n_real) ** 2 <= 1.3
                                  or self.rows_total * self.npad * self.npad <= RESIDENT_PAD_WORK))
        self.pad = self.resident and self.npad != n_real
        self.Xb = None
        if self.pad:
            self.X, self.Xb = _padded2d(B, k, cplx, self.npad, dev)
            X = self.X
        else:
            self.X = X = empty2d(B, k, cplx, dev)
        if x is None:
            X.zero_()                                              # default x = zeros (lasso.py:73-74)
        else:
            ops.scale(rview(x), rview(X), cwidth=cw, colscale=s)

        # ---- Gram matrix, step 1/L, yAh (lasso.py:276-289, 306-321)
        self.AH = AH = ops.make_rhs(Anr, True, True) if cplx else Anr    # NT operand of  . A^H
        G = empty2d(k, k, cplx, dev)
        self.rowvec = None
        if full_mask:
            self.rowvec = ops.row_sums(mask)                       # sum(mask, -1): alpha per problem
            if group is not None:
                mean = _global_mask_mean(mask, B, group)
            else:
                mean = ops.col_sums(mask, 1.0 / B)
            Amean = empty2d(k, f, cplx, dev)
            ops.scale(Anr, rview(Amean), cwidth=cw, colscale=mean)
            ops.gemm_nt(rview(Amean), AH, ops.epilogue(ops.EPI_STORE, rview(G)))
        else:
            ops.gemm_nt(Anr, AH, ops.epilogue(ops.EPI_STORE, rview(G)))
        self.step = torch.empty(1, dtype=torch.float64, device=dev)
        # threshold step * alpha (lasso.py:287) is formed once here unless alpha is per problem (full mask)
        self.thr = None if full_mask else ops.vector(k, dev)
        ops.gershgorin_step(rview(G), cplx, self.step, alpha_scaled=self.alpha_vec, thr_out=self.thr)
        if self.pad:
            self.yAh, self.Cb = _padded2d(B, k, cplx, self.npad, dev)
            yAh = self.yAh
        else:
            self.yAh = yAh = empty2d(B, k, cplx, dev)
        if full_mask:
            self.T = T = empty2d(B, f, cplx, dev)                  # also the per-iteration [B, f] temporary
            ops.mask_mul(yr, mask, rview(T), cwidth=cw)
            ops.gemm_nt(rview(T), AH, ops.epilogue(ops.EPI_STORE, rview(yAh)))
            self.A_rhs = ops.make_rhs(Anr, cplx, False)            # NT operand of  w . A
            # widths the fused back-to-back kernel covers: both masked GEMMs of an iteration in one launch
            self.b2b = bool(USE_B2B and not self.tf32 and ops.gemm_b2b_masked_supported(k * cw))
            if self.b2b or self.tf32:
                self.T = None                                      # only the set-up needed the [B, f] temporary
        else:
            ops.gemm_nt(yr, AH, ops.epilogue(ops.EPI_STORE, rview(yAh)))
            # fold the gradient step into the operands:  w + (yAh - w G)/L  =  yAh/L + w (I - G/L)
            Q = empty2d(k, k, cplx, dev)
            ops.lasso_q(rview(G), cplx, self.step, rview(Q))
            self.Q_rhs = ops.make_rhs(rview(Q), cplx, False)       # NT operand of  w . Q
            if self.pad:
                # zero rows / columns for the padding: it stays 0 through threshold and extrapolation
                Qp = torch.zeros((self.npad, self.npad), dtype=torch.float64, device=dev)
                Qp[:n_real, :n_real].copy_(self.Q_rhs)
                self.Q_pad = Qp
                nv = self.npad // cw
                self.thr_pad = torch.zeros(nv, dtype=torch.float64, device=dev)
                self.thr_pad[:k].copy_(self.thr)
                self.tol_pad = torch.ones(nv, dtype=torch.float64, device=dev)
                self.tol_pad[:k].copy_(self.tol_vec)
            ops.scale_scalar(rview(yAh), self.step, rview(yAh))    # yAh <- yAh / L
        if self.tf32 and full_mask:
            # masked iteration on the tcgen05 tensor cores (lasso.py:259-271): t = (w A) * mask as a TF32 pair through
            # HBM, then t A^H, then the FP64 threshold / momentum pass
            n_real = k * cw
            if n_real % 2 != 0:
                raise NotImplementedError("precision='tf32x3' with a mask needs an even k for real data; use "
                                          "precision='fp64'")
            self.M32 = ops.to_f32(mask)
            self.Ar_hi, self.Ar_lo = ops.split_tf32(self.A_rhs)    # [f cw, k cw]: NT operand of w . A
            self.AH_hi, self.AH_lo = ops.split_tf32(AH)            # [k cw, f cw]: NT operand of t . A^H
            self.T_hi, self.T_lo = ops.empty_f32(B, f * cw, dev), ops.empty_f32(B, f * cw, dev)
            self.P = ops.empty_f32(B, n_real, dev)
        elif self.tf32:
            n_real = k * cw
            if n_real % 32 != 0 or n_real > 256:
                raise NotImplementedError("precision='tf32x3' covers the unmasked iteration with k (2k for complex "
                                          "data) a multiple of 32 up to 256; use precision='fp64'")
            self.Q_hi, self.Q_lo = ops.split_tf32(self.Q_rhs)
            self.P = ops.empty_f32(B, n_real, dev)

        # ---- iteration state
        self.checks = checks = tol > 0.0
        self.latch = torch.zeros(1, dtype=torch.int32, device=dev) if checks else None
        self.scratch = torch.zeros(2, dtype=torch.int32, device=dev) if checks else None
        if self.tf32:
            self.W_hi, self.W_lo = ops.split_tf32(rview(X))        # w0 = x0 as a TF32 pair, updated in place
            self.W = [X, X]
        elif self.pad:
            w0, wb0 = _padded2d(B, k, cplx, self.npad, dev)        # updated in place by the resident kernel
            self.W, self.Wb = [w0, w0], [wb0, wb0]
            if rule == 'acc_ista':                                 # its last iteration may run per-iteration (finish)
                w1, wb1 = _padded2d(B, k, cplx, self.npad, dev)
                self.W, self.Wb = [w0, w1], [wb0, wb1]
            w0.copy_(X)
        else:
            self.W = [empty2d(B, k, cplx, dev), empty2d(B, k, cplx, dev)]
            self.W[0].copy_(X)
        self.wi = 0                                                # W[wi] holds the current extrapolated point
        self.mom = _momentum_schedule(rule, maxiter)
        # acc_ista returns the *previous* iterate on exhaustion (lasso.py:357,385): its last iteration only
        # matters if it is a checking one, and then only when the check passes.
        self.n_inplace = maxiter - 1 if (rule == 'acc_ista' and maxiter > 0) else maxiter
        self.stopped = False

    def _launch(self, i, out_x):
        W, cw, latch = self.W, self.cw, self.latch
        check = self.checks and i % 10 == 0
        if self.tf32 and self.full_mask:
            epi = ops.epilogue(ops.EPI_PROX, out_x, cwidth=cw, other=rview(self.yAh), prev=rview(self.X),
                               colvec=self.alpha_vec, colvec2=self.tol_vec, rowvec=self.rowvec, step=self.step,
                               momentum=self.mom[i], shrink=self.shrink, check=check, latch=latch,
                               scratch=self.scratch, latch_value=i + 1)
            ops.gemm_nt_mask_tf32x3(self.W_hi, self.W_lo, self.Ar_hi, self.Ar_lo, self.M32, F=(self.T_hi, self.T_lo),
                                    cwidth=cw, skip=latch)
            ops.gemm_nt_tf32x3(self.T_hi, self.T_lo, self.AH_hi, self.AH_lo, self.P, skip=latch)
            ops.prox_apply(self.P, epi, self.W_hi, self.W_lo, skip=latch)
            self._exchange_latch(check, i + 1)
            return
        if self.tf32:
            epi = ops.epilogue(ops.EPI_PROXQ, out_x, cwidth=cw, other=rview(self.yAh), prev=rview(self.X),
                               colvec=self.thr, colvec2=self.tol_vec, flags=ops.EPI_FLAG_COLVEC_IS_THRESHOLD,
                               momentum=self.mom[i], shrink=self.shrink, check=check, latch=latch,
                               scratch=self.scratch, latch_value=i + 1)
            ops.gemm_nt_tf32x3(self.W_hi, self.W_lo, self.Q_hi, self.Q_lo, self.P, skip=latch)
            ops.proxq_apply(self.P, epi, self.W_hi, self.W_lo, skip=latch)
            self._exchange_latch(check, i + 1)
            return
        wi = self.wi
        self.wi = 1 - wi
        if not self.full_mask:
            epi = ops.epilogue(ops.EPI_PROXQ, out_x, cwidth=cw, out2=rview(W[1 - wi]), other=rview(self.yAh),
                               prev=rview(self.X), colvec=self.thr, colvec2=self.tol_vec,
                               flags=ops.EPI_FLAG_COLVEC_IS_THRESHOLD, momentum=self.mom[i], shrink=self.shrink,
                               check=check, latch=latch, scratch=self.scratch, latch_value=i + 1)
            ops.gemm_nt(rview(W[wi]), self.Q_rhs, epi, skip=latch)
            self._exchange_latch(check, i + 1)
            return
        epi = ops.epilogue(ops.EPI_PROX, out_x, cwidth=cw, out2=rview(W[1 - wi]), x=rview(W[wi]),
                           other=rview(self.yAh), prev=rview(self.X),
                           colvec=self.alpha_vec, colvec2=self.tol_vec,
                           rowvec=self.rowvec, step=self.step, momentum=self.mom[i], shrink=self.shrink,
                           check=check, latch=latch, scratch=self.scratch, latch_value=i + 1, mask=self.mask)
        if self.b2b:
            # ((w A) * M) A^H in one kernel, the [B, f] intermediate stays on chip (lasso.py:259-271)
            ops.gemm_b2b_masked(rview(W[wi]), self.A_rhs, epi, skip=latch)
        else:
            ops.gemm_nt(rview(W[wi]), self.A_rhs,
                        ops.epilogue(ops.EPI_STORE_MASK, rview(self.T), cwidth=cw, mask=self.mask), skip=latch)
            ops.gemm_nt(rview(self.T), self.AH, epi, skip=latch)
        self._exchange_latch(check, i + 1)

    def _exchange_latch(self, check, value):
        """The latch fires only if every shard passed the test (reference: one max over the whole batch, lasso.py:293):
        MIN over the ranks.  A rank without rows (a minibatch of dictionary learning may leave it none) launches no
        kernel that could vote, so it passes by setting its latch itself."""
        if check and self.group is not None:
            if self.B == 0:
                self.latch.fill_(value)
            comm.all_reduce_min_i32(self.latch, self.group)

    def _launch_resident(self, i0, i1):
        """Iterations i0 <= i < i1 in one launch; the convergence test may only sit on the last one."""
        latch = self.latch
        check = self.checks and (i1 - 1) % 10 == 0
        if self.pad:
            X, W, C, Q = self.Xb, self.Wb[self.wi], self.Cb, self.Q_pad
            thr, tolv = self.thr_pad, self.tol_pad
        else:
            X, W, C, Q = rview(self.X), rview(self.W[self.wi]), rview(self.yAh), self.Q_rhs
            thr, tolv = self.thr, self.tol_vec
        epi = ops.epilogue(ops.EPI_PROXQ, X, cwidth=self.cw, x=W, other=C, colvec=thr, colvec2=tolv,
                           flags=ops.EPI_FLAG_COLVEC_IS_THRESHOLD, shrink=self.shrink, check=check, latch=latch,
                           scratch=self.scratch, latch_value=i1)
        ops.lasso_resident(Q, self.B, epi, self.mom[i0:i1], skip=latch)
        self._exchange_latch(check, i1)

    def _poll(self, last_done):
        """Host read of the latch, on the same schedule in both loops: right after checking iteration ``last_done``
        has been enqueued, every POLL_EVERY iterations.  (With a group every rank must leave the loop at the same
        iteration, whatever kernel path its shard size selected: the latch is MIN-all-reduced on checking iterations,
        so ranks that poll after the same iterations read the same value.)"""
        if self.checks and last_done > 0 and last_done % POLL_EVERY == 0 and int(self.latch.item()) != 0:
            self.stopped = True
        return self.stopped

    def iterate(self, begin, end):
        """Enqueue iterations ``begin <= i < end`` (in place on X)."""
        end = min(end, self.n_inplace)
        if self.stopped:
            return
        if self.resident:
            i = begin
            while i < end:
                stop = min(end, i + ops.RESIDENT_MAX_ITERS)
                if self.checks:
                    stop = min(stop, (i + 9) // 10 * 10 + 1)       # a launch ends on the next checking iteration
                if stop - i <= RESIDENT_MIN_ITERS and not self.pad and self.rows_total * self.npad ** 2 > RESIDENT_PAD_WORK * 3:
                    # a resident launch reads and writes the whole state once (~0.27 ms at C2) whatever its length:
                    # short launches of big batches are cheaper one iteration at a time
                    for j in range(i, stop):
                        self._launch(j, rview(self.X))
                else:
                    self._launch_resident(i, stop)
                i = stop
                if self._poll(stop - 1):
                    break
            return
        for i in range(begin, end):
            self._launch(i, rview(self.X))
            if self._poll(i):
                break

    def finish(self, out=None):
        """x / s (lasso.py:189) into ``out``; returns a ``LassoState``."""
        final, maxiter = self.X, self.maxiter
        if (self.rule == 'acc_ista' and maxiter > 0 and not self.stopped and self.checks
                and (maxiter - 1) % 10 == 0):
            XL = empty2d(self.B, self.k, self.cplx, self.X.device)
            self._launch(maxiter - 1, rview(XL))
            if int(self.latch.item()) == maxiter:
                final = XL
        if out is None:
            out = empty2d(self.B, self.k, self.cplx, self.X.device)
        ops.scale(rview(final), rview(out), cwidth=self.cw, colscale=self.s, invert_col=True)
        return LassoState(out, self.latch, maxiter)


def _global_mask_mean(mask, local_rows, group):
    """mean over the whole (sharded) batch of the mask, lasso.py:300-303: one [f] all-reduce at set-up."""
    dist = torch.distributed
    sums = ops.col_sums(mask, 1.0)
    rows = torch.tensor([float(local_rows)], dtype=torch.float64, device=mask.device)
    comm.all_reduce_sum(sums, group)
    comm.all_reduce_sum(rows, group)
    f = sums.numel()
    mean = torch.empty_like(sums)
    ops.scale(sums.view(1, f), mean.view(1, f), rowscale=rows, invert_row=True)
    return mean
