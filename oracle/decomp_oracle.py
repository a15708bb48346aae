"""CPU oracle for the deComP hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-numpy restatement of the reference's algorithm for the three solvers on the
hot path (NMF multiplicative update, batched ISTA/FISTA Lasso, online dictionary
learning).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the product
package ``decomp_b200`` never does (it fails loudly when the CUDA library is
missing instead of falling back to anything here).

Parity status: PINNED.  ``tools/make_golden.py`` imports the unmodified reference
from ``/root/reference`` in the build container, runs it on seeded inputs and
stores iteration counts and factors under ``tests/golden/``;
``tests/test_oracle_golden.py`` replays the same inputs through this module and
requires the same iteration counts and factors (<= 1e-12 relative).  The
reference's own literal known answers (soft-threshold vectors,
``tests/test_lasso.py:15-56``) are checked there too.

Every function cites the reference lines it follows (paths relative to
``/root/reference``).  The arithmetic keeps the reference's operation order so
that the numpy results agree to rounding; the code organisation is our own.
"""
import numpy as np

EPS = 1.0e-15  # the reference's _JITTER (nmf.py:13, grads.py:4, lasso.py:16, dictionary_learning.py:9)

LASSO_RULES = ('ista', 'fista', 'acc_ista')


# --------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------
def row_energy(U):
    """sum_j |U[..., j]|^2 with keepdims (decomp/utils/normalize.py:6-9, 17-20)."""
    if np.iscomplexobj(U):
        return np.sum(np.real(np.conj(U) * U), axis=-1, keepdims=True)
    return np.sum(U * U, axis=-1, keepdims=True)


def unit_rows(U):
    """Rows scaled to exactly unit L2 norm (decomp/utils/normalize.py:13-21 ``l2_strict``)."""
    return U / np.sqrt(row_energy(U))


def shrink_rows(U):
    """Rows scaled to norm <= 1 (decomp/utils/normalize.py:2-10 ``l2``)."""
    return U / np.sqrt(np.maximum(row_energy(U), 1.0))


def gershgorin(X):
    """max_j sum_i |X_ij| with the last axis kept (decomp/math_utils/eigen.py:9-20)."""
    return np.max(np.sum(np.abs(X), axis=-2), axis=-1, keepdims=True)


def shrink_real(z, t):
    """decomp/lasso.py:192-207."""
    s = np.sign(z)
    return np.maximum(np.abs(z) - t, 0.0) * s


def shrink_complex(z, t):
    """decomp/lasso.py:210-225."""
    r = np.abs(z)
    phase = z / (r + EPS)
    return np.maximum(r - t, 0.0) * phase


def shrink_positive(z, t):
    """decomp/lasso.py:228-241."""
    return np.maximum(z - t, 0.0)


# --------------------------------------------------------------------------------------
# NMF, full-batch multiplicative update, Gaussian (l2) and Poisson (kl) likelihoods
# --------------------------------------------------------------------------------------
def _mu_parts_x(y, x, d, mask, kl):
    """Positive / negative gradient parts for x (decomp/nmf_methods/grads.py:108-115, 142-149)."""
    if not kl:
        if mask is None:
            f = x.dot(d)
            return y.dot(d.T), f.dot(d.T)
        f = x.dot(d) * mask
        ym = y * mask
        return ym.dot(d.T), f.dot(d.T)
    f = x.dot(d) + EPS
    if mask is None:
        return (y / f).dot(d.T), d.T.sum(axis=0, keepdims=True)
    ym = y * mask
    return (ym / f).dot(d.T), mask.dot(d.T)


def _mu_parts_d(y, x, d, mask, kl):
    """Positive / negative gradient parts for D (decomp/nmf_methods/grads.py:117-125, 151-160)."""
    if not kl:
        if mask is None:
            f = x.dot(d)
            return x.T.dot(y), x.T.dot(f)
        f = x.dot(d) * mask
        ym = y * mask
        return x.T.dot(ym), x.T.dot(f)
    f = x.dot(d) + EPS
    if mask is None:
        return x.T.dot(y / f), x.T.sum(axis=1, keepdims=True)
    ym = y * mask
    return x.T.dot(ym / f), x.T.dot(mask)


def nmf_mu(y, D, x=None, tol=1.0e-3, maxiter=1000, likelihood='l2', mask=None):
    """Full-batch NMF-MU: decomp/nmf.py:52-78 -> decomp/nmf_methods/batch_mu.py:8-26.

    Returns ``(it, D, x)`` with the reference's conventions: at most ``maxiter - 1``
    sweeps, ``it == maxiter`` on exhaustion, default ``x`` is all ones.
    """
    kl = likelihood in ('kl', 'poisson')
    if x is None:
        x = np.ones((y.shape[0], D.shape[0]), dtype=y.dtype)
    D = unit_rows(D)
    for it in range(1, maxiter):
        pos, neg = _mu_parts_x(y, x, D, mask, kl)
        x = x * np.maximum(pos, 0.0) / np.maximum(neg, EPS)       # grads.py:77-84
        pos, neg = _mu_parts_d(y, x, D, mask, kl)
        D_next = D * np.maximum(pos, 0.0) / np.maximum(neg, EPS)  # grads.py:86-93
        D_next = unit_rows(D_next)
        if np.max(np.abs(D - D_next)) < tol:
            return it, D_next, x
        D = D_next
    return maxiter, D, x


# --------------------------------------------------------------------------------------
# NMF, minibatch multiplicative updates (decomp/nmf.py:82-111 -> nmf_methods/serizel.py, kasai.py)
# --------------------------------------------------------------------------------------
def _mu_step(v, pos, neg):
    return v * np.maximum(pos, 0.0) / np.maximum(neg, EPS)


def nmf_minibatch(y, D, x=None, tol=1.0e-3, minibatch=10, maxiter=1000, method='asg-mu', likelihood='l2', mask=None,
                  random_seed=None, forget_rate=0.5, alpha=1.0, beta=0.5):
    """Minibatch NMF drivers of the reference, with its quirks: 'gsg-mu' runs the 'asg-mu' loop
    (serizel.py:23-25), and on convergence the *previous* D is returned (serizel.py:58, kasai.py:81)."""
    kl = likelihood in ('kl', 'poisson')
    if x is None:
        x = np.ones((y.shape[0], D.shape[0]), dtype=y.dtype)
    D = unit_rows(D)                                              # nmf.py:70
    ys, xs = _Rows(y, minibatch), _Rows(x, minibatch)
    ms = _Rows(mask, minibatch) if mask is not None else None
    rng = np.random.RandomState(random_seed)
    order = np.arange(len(y))
    n_loop = len(y) // minibatch

    def permute():
        rng.shuffle(order)
        ys.permute(order)
        xs.permute(order)
        if ms is not None:
            ms.permute(order)

    def batches():
        mc = ms.chunks() if ms is not None else None
        for y_mb, x_mb in zip(ys.chunks(), xs.chunks()):
            yield y_mb, x_mb, (next(mc) if mc is not None else None)

    def update_x(y_mb, x_mb, m_mb, D):
        pos, neg = _mu_parts_x(y_mb, x_mb, D, m_mb, kl)
        x_mb[:] = _mu_step(x_mb, pos, neg)

    if method in ('asg-mu', 'gsg-mu'):                            # serizel.py:36-61
        for it in range(1, maxiter):
            permute()
            for y_mb, x_mb, m_mb in batches():
                update_x(y_mb, x_mb, m_mb, D)
                pos, neg = _mu_parts_d(y_mb, x_mb, D, m_mb, kl)
                D_new = unit_rows(_mu_step(D, pos, neg))
                if np.max(np.abs(D - D_new)) < tol:
                    return it, D, xs.restored()
                D = D_new
        return maxiter, D, xs.restored()

    if method in ('asag-mu', 'gsag-mu'):                          # serizel.py:92-165
        every_batch = method == 'asag-mu'
        for it in range(1, maxiter):
            permute()
            pos_sum, neg_sum = np.zeros_like(D), np.zeros_like(D)
            for y_mb, x_mb, m_mb in batches():
                update_x(y_mb, x_mb, m_mb, D)
                pos, neg = _mu_parts_d(y_mb, x_mb, D, m_mb, kl)
                pos_sum = (1.0 - forget_rate) * pos_sum + forget_rate * pos
                neg_sum = (1.0 - forget_rate) * neg_sum + forget_rate * neg
                if every_batch:
                    D_new = unit_rows(_mu_step(D, pos_sum, neg_sum))
                    if np.max(np.abs(D - D_new)) < tol:
                        return it, D, xs.restored()
                    D = D_new
            if not every_batch:
                D_new = unit_rows(_mu_step(D, pos_sum, neg_sum))
                if np.max(np.abs(D - D_new)) < tol:
                    return it, D, xs.restored()
                D = D_new
        return maxiter, D, xs.restored()

    if method in ('svrmu', 'svrmu-acc'):                          # kasai.py:10-88
        if method == 'svrmu':
            inner = 1
        else:
            F, K = D.shape
            N = x.shape[0]
            inner = int(np.maximum(beta * F * (3 * K + 2 * N) / (3 * F * N + 2 * K), 1.0))
        permute()                                                  # once, before the first epoch
        pos_prev = np.zeros((n_loop,) + D.shape, dtype=D.dtype)
        neg_prev = np.zeros((n_loop,) + D.shape, dtype=D.dtype)
        for it in range(1, maxiter):
            pos_full, neg_full = np.zeros_like(D), np.zeros_like(D)
            for y_mb, x_mb, m_mb in batches():
                pos, neg = _mu_parts_d(y_mb, x_mb, D, m_mb, kl)
                pos_full += pos
                neg_full += neg
            pos_full /= n_loop
            neg_full /= n_loop
            for b, (y_mb, x_mb, m_mb) in enumerate(batches()):
                for _ in range(inner):
                    update_x(y_mb, x_mb, m_mb, D)
                pos, neg = _mu_parts_d(y_mb, x_mb, D, m_mb, kl)
                P = pos + neg_prev[b] + pos_full
                Q = neg + pos_prev[b] + neg_full
                D_new = D * ((1.0 - alpha) + alpha * P / np.maximum(Q, EPS))
                D_new = unit_rows(np.maximum(D_new, 0.0))
                if np.max(np.abs(D - D_new)) < tol:
                    return it, D, xs.restored()
                D = D_new
                pos_prev[b] = pos
                neg_prev[b] = neg
        return maxiter, D, xs.restored()
    raise NotImplementedError('NMF with {} algorithm is not yet implemented.'.format(method))


# --------------------------------------------------------------------------------------
# batched Lasso: ISTA / FISTA / accelerated ISTA, optional non-negativity and masks
# --------------------------------------------------------------------------------------
def _batch_mean(mask):
    """Mean over every axis but the last (decomp/lasso.py:300-303)."""
    for _ in range(mask.ndim - 1):
        mask = np.mean(mask, 0)
    return mask


def lasso(y, A, alpha, x=None, tol=1.0e-3, method='ista', maxiter=1000, mask=None):
    """decomp/lasso.py:97-189 (``solve_fastpath``) plus the default ``x`` of ``solve`` (:73-74).

    ``method`` is one of ista / fista / acc_ista, optionally suffixed ``_pos``.
    """
    if x is None:
        x = np.zeros(y.shape[:-1] + (A.shape[0],), dtype=y.dtype)
    positive = method.endswith('_pos')
    rule = method[:-4] if positive else method
    if rule not in LASSO_RULES:
        raise ValueError('oracle covers ' + str(LASSO_RULES) + ' only, given ' + method)

    if mask is not None and mask.ndim == 1:            # lasso.py:120-122
        y = y * mask
        A = A * mask
    # unit-diagonal scaling of A (lasso.py:124-131)
    if np.iscomplexobj(A):
        scale = np.sqrt(np.sum(np.real(np.conj(A) * A), axis=-1))
    else:
        scale = np.sqrt(np.sum(np.square(A), axis=-1))
    A = A / np.expand_dims(scale, axis=-1)
    alpha = alpha / scale
    tol = tol * scale
    x = x * scale

    full_mask = mask is not None and mask.ndim > 1
    if full_mask:
        alpha = alpha * np.sum(mask, axis=-1, keepdims=True)      # lasso.py:163
    elif mask is not None:
        alpha = alpha * np.sum(mask, axis=-1)                      # lasso.py:136
    else:
        alpha = alpha * A.shape[-1]                                # lasso.py:138

    if positive:
        Ah, shrink = A.T, shrink_positive
    elif not np.iscomplexobj(A):
        Ah, shrink = A.T, shrink_real
    else:
        Ah, shrink = np.conj(A.T), shrink_complex

    if full_mask:                                                  # lasso.py:317-321, 429-433
        gram = np.dot(A * _batch_mean(mask), Ah)
    else:                                                          # lasso.py:285-289, 399-403
        gram = np.dot(A, Ah)
    step = 1.0 / gershgorin(gram)
    thresh = step * alpha
    if full_mask:
        yAh = np.tensordot(y * mask, Ah, axes=1)

        def prox_grad(p):                                          # lasso.py:259-271
            resid = yAh - np.tensordot(np.tensordot(p, A, axes=1) * mask, Ah, axes=1)
            return shrink(p + step * resid, thresh)
    else:
        yAh = np.tensordot(y, Ah, axes=1)

        def prox_grad(p):                                          # lasso.py:244-256
            resid = yAh - np.tensordot(p, gram, axes=1)
            return shrink(p + step * resid, thresh)

    def settled(a, b):
        return np.max(np.abs(a - b) - tol) < 0.0

    it = maxiter - 1
    if rule == 'ista':                                             # lasso.py:291-297, 323-328
        cur = x
        for i in range(maxiter):
            nxt = prox_grad(cur)
            if i % 10 == 0 and settled(nxt, cur):
                cur, it = nxt, i
                break
            cur = nxt
        out = cur
    elif rule == 'fista':                                          # lasso.py:405-415, 435-445
        cur, probe, beta = x, x, 1.0
        for i in range(maxiter):
            nxt = prox_grad(probe)
            if i % 10 == 0 and settled(nxt, cur):
                cur, it = nxt, i
                break
            beta_next = 0.5 * (1.0 + np.sqrt(1.0 + 4.0 * beta * beta))
            probe = nxt + (beta - 1.0) / beta_next * (nxt - cur)
            cur = nxt
            beta = beta_next
        out = cur
    else:                                                          # acc_ista, lasso.py:348-357, 377-385
        probe, nxt, prev = x, x, x
        out = None
        for i in range(maxiter):
            prev = nxt
            nxt = prox_grad(probe)
            probe = nxt + i / (i + 3) * (nxt - prev)
            if i % 10 == 0 and settled(nxt, prev):
                out, it = nxt, i
                break
        if out is None:
            out = prev      # the reference returns the *previous* iterate on exhaustion (:357, :385)
    return it, out / scale                                         # lasso.py:189


# --------------------------------------------------------------------------------------
# online dictionary learning (Mairal block coordinate descent)
# --------------------------------------------------------------------------------------
class _Rows(object):
    """Row-permuted view that remembers how to undo itself (decomp/utils/data.py:124-156)."""

    def __init__(self, array, step):
        if len(array) < step:                                      # data.py:79-82
            raise ValueError('Minibatch size should be smaller than the total size.')
        self.data = array
        self.step = step
        self.origin = np.arange(len(array))

    def permute(self, index):
        self.data = self.data[index]
        self.origin = self.origin[index]

    def chunks(self):
        for r in range(len(self.data) // self.step):               # tail rows are skipped (data.py:101-103)
            yield self.data[r * self.step:(r + 1) * self.step]

    def restored(self):
        return self.data[self.origin.argsort()]


def dictionary_learning(y, D, alpha, x=None, tol=1.0e-3, minibatch=None, maxiter=1000,
                        lasso_method='cd', lasso_iter=10, lasso_tol=1.0e-5, mask=None,
                        random_seed=None):
    """decomp/dictionary_learning.py:12-231 (``solve`` -> ``solve_cd`` / ``solve_cd_mask``)."""
    if x is None:
        x = np.ones((y.shape[0], D.shape[0]), dtype=D.dtype)       # :58-59
    if minibatch is None:
        raise NotImplementedError('Only online methods are implemented. minibatch is required.')
    rng = np.random.RandomState(random_seed)
    ys, xs = _Rows(y, minibatch), _Rows(x, minibatch)
    ms = _Rows(mask, minibatch) if mask is not None else None
    n_atoms, n_feat = D.shape
    is_complex = np.iscomplexobj(y)
    order = np.arange(len(y))
    if ms is None:
        S = np.zeros((n_atoms, n_atoms), dtype=y.dtype)            # :122
    else:
        S = np.zeros((n_atoms, n_feat, n_atoms), dtype=y.dtype)    # :179
    T = np.zeros((n_atoms, n_feat), dtype=y.dtype)
    D = unit_rows(D)
    seen = 0
    for it in range(1, maxiter):
        rng.shuffle(order)                                         # cumulative permutation (:131-133)
        ys.permute(order)
        xs.permute(order)
        if ms is not None:
            ms.permute(order)
        mask_chunks = ms.chunks() if ms is not None else None
        for y_mb, x_mb in zip(ys.chunks(), xs.chunks()):
            m_mb = next(mask_chunks) if mask_chunks is not None else None
            _, code = lasso(y_mb, D, alpha, x=x_mb, tol=lasso_tol, maxiter=lasso_iter,
                            method=lasso_method, mask=m_mb)        # :137-139, :195-198
            x_mb[...] = code
            theta = seen * minibatch + 1.0                         # :143-144
            forget = (theta - minibatch) / theta
            xh = np.conj(x_mb.T) if is_complex else x_mb.T
            D_next = D.copy()
            if m_mb is None:
                S = forget * S + np.dot(xh, x_mb)                  # :151-152
                T = forget * T + np.dot(xh, y_mb)
                for a in range(n_atoms):                           # Gauss-Seidel sweep (:155-159)
                    u = (T[a] - np.dot(S[a], D_next)) / (S[a, a] + EPS) + D_next[a]
                    D_next[a] = shrink_rows(u)
            else:
                S = forget * S + np.tensordot(
                    xh, np.expand_dims(x_mb, -2) * np.expand_dims(m_mb, -1), axes=1)   # :210-213
                T = forget * T + np.dot(xh, y_mb * m_mb)           # :214
                for a in range(n_atoms):                           # Jacobi update against the old D (:217-222)
                    SaD = np.einsum('jk,kj->j', S[a], D)
                    Saa = np.sum(S[a, :, a] + EPS)
                    u = (T[a] - SaD) / Saa + D_next[a]
                    D_next[a] = shrink_rows(u)
            if np.max(np.abs(D - D_next)) < tol:                   # :161-162, :224-225
                return it, D_next, xs.restored()
            D = D_next
            seen += 1
    return maxiter, D, xs.restored()


# --------------------------------------------------------------------------------------
# objectives used by the parity tests (tests/test_nmf.py:29-39, tests/test_lasso.py:127-134)
# --------------------------------------------------------------------------------------
def nmf_objective(y, x, D, mask=None):
    r = np.square(y - np.dot(x, unit_rows(D)))
    if mask is not None:
        r = r * mask
    return 0.5 * np.sum(r)


def lasso_objective(y, A, x, alpha, mask=None):
    if mask is None:
        mask = np.ones(y.shape, dtype=np.zeros(1, y.dtype).real.dtype)
    elif mask.ndim == 1:
        mask = np.ones(y.shape, dtype=mask.dtype) * mask
    a = alpha * np.sum(mask, axis=-1, keepdims=True)
    loss = np.sum(0.5 / a * np.square(np.abs(y - np.tensordot(x, A, axes=1))) * mask)
    return loss + np.sum(np.abs(x))
