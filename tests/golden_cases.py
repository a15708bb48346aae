"""Seeded input generators shared by tools/make_golden.py, the oracle tests and the GPU
parity tests.  Data recipes mirror the reference's own tests (cited per function) so the
parity tests read like the reference's; sizes are kept small enough that the golden
outputs stay a few hundred kB.  ``numpy.random.RandomState`` is the legacy generator
whose streams are frozen across numpy versions.
"""
from collections import OrderedDict

import numpy as np


# ------------------------------------------------------------------------------- NMF
def _nmf_data(n, f, k, seed, likelihood='l2', reference_order=True):
    rng = np.random.RandomState(seed)
    if reference_order:      # /root/reference/tests/test_nmf.py:60-69
        Dt = np.maximum(rng.randn(k, f), 0.0)
        xt = np.maximum(rng.randn(n, k), 0.0)
    else:                    # SURVEY.md section 8(d), config C1 ordering
        xt = np.maximum(rng.randn(n, k), 0.0)
        Dt = np.maximum(rng.randn(k, f), 0.0)
    y = np.dot(xt, Dt)
    noise = rng.randn(*y.shape)
    y = y + (np.abs(noise) if likelihood == 'kl' else noise) * 0.1
    D0 = np.maximum(Dt + rng.randn(k, f) * 0.3, 0.1)
    mask = np.rint(rng.uniform(0.3, 1, size=y.size)).reshape(y.shape)
    return y, D0, mask


def nmf_cases():
    cases = OrderedDict()
    for lik in ('l2', 'kl'):
        y, D0, mask = _nmf_data(101, 20, 3, 0, lik)
        cases['test_%s' % lik] = dict(y=y, D=D0, mask=None, tol=1.0e-6, maxiter=3000, likelihood=lik)
        cases['test_%s_mask' % lik] = dict(y=y, D=D0, mask=mask, tol=1.0e-6, maxiter=3000, likelihood=lik)
    # BASELINE config 1: Y 1000x200, k=20, exactly 100 sweeps
    y, D0, mask = _nmf_data(1000, 200, 20, 0, 'l2', reference_order=False)
    cases['c1_l2'] = dict(y=y, D=D0, mask=None, tol=0.0, maxiter=101, likelihood='l2')
    cases['c1_l2_mask'] = dict(y=y, D=D0, mask=mask, tol=0.0, maxiter=101, likelihood='l2')
    # ragged sizes: nothing divides a tile
    y, D0, mask = _nmf_data(257, 67, 7, 3, 'l2')
    cases['ragged_l2'] = dict(y=y, D=D0, mask=None, tol=0.0, maxiter=41, likelihood='l2')
    cases['ragged_l2_mask'] = dict(y=y, D=D0, mask=mask, tol=0.0, maxiter=41, likelihood='l2')
    y, D0, mask = _nmf_data(257, 67, 7, 3, 'kl')
    cases['ragged_kl_mask'] = dict(y=y, D=D0, mask=mask, tol=0.0, maxiter=41, likelihood='kl')
    return cases


def nmf_minibatch_cases():
    """Minibatch NMF drivers (/root/reference/tests/test_nmf.py:105-152 uses the same data recipe)."""
    cases = OrderedDict()
    y, D0, mask = _nmf_data(101, 20, 3, 0, 'l2')
    ykl, D0kl, _ = _nmf_data(101, 20, 3, 0, 'kl')
    for method in ('asg-mu', 'gsg-mu', 'asag-mu', 'gsag-mu', 'svrmu', 'svrmu-acc'):
        tag = method.replace('-', '_')
        cases['%s_l2' % tag] = dict(y=y, D=D0, mask=None, tol=0.0, maxiter=8, minibatch=10, method=method,
                                    likelihood='l2', random_seed=0)
        cases['%s_l2_mask' % tag] = dict(y=y, D=D0, mask=mask, tol=0.0, maxiter=8, minibatch=10, method=method,
                                         likelihood='l2', random_seed=1)
    cases['asg_mu_kl_mask'] = dict(y=ykl, D=D0kl, mask=mask, tol=0.0, maxiter=6, minibatch=25, method='asg-mu',
                                   likelihood='kl', random_seed=2)
    cases['svrmu_kl'] = dict(y=ykl, D=D0kl, mask=None, tol=0.0, maxiter=6, minibatch=25, method='svrmu',
                             likelihood='kl', random_seed=2)
    cases['asag_mu_conv'] = dict(y=y, D=D0, mask=None, tol=1.0e-3, maxiter=400, minibatch=20, method='asag-mu',
                                 likelihood='l2', random_seed=3)
    y2, D02, mask2 = _nmf_data(257, 67, 7, 3, 'l2')
    cases['gsag_mu_ragged_mask'] = dict(y=y2, D=D02, mask=mask2, tol=0.0, maxiter=5, minibatch=50, method='gsag-mu',
                                        likelihood='l2', random_seed=4)
    cases['svrmu_acc_ragged'] = dict(y=y2, D=D02, mask=None, tol=0.0, maxiter=5, minibatch=50, method='svrmu-acc',
                                     likelihood='l2', random_seed=4)
    return cases


# ------------------------------------------------------------------------------- Lasso
def _lasso_data(batch_shape, k, f, seed, complex_=False, dtype=None, positive=False):
    """/root/reference/tests/test_lasso.py:143-150, 223-250 (vector / matrix / tensor set-ups)."""
    rng = np.random.RandomState(seed)

    def randn(*shape):
        if complex_:
            return rng.randn(*shape) + rng.randn(*shape) * 1.0j
        return rng.randn(*shape)

    nb = int(np.prod(batch_shape)) if batch_shape else 1
    A = randn(k, f)
    if positive:
        x_true = np.maximum(randn(nb * k), 0.0)
    else:
        x_true = randn(nb * k) * np.rint(rng.uniform(size=nb * k))
    x_true = x_true.reshape(tuple(batch_shape) + (k,))
    y = np.tensordot(x_true, A, axes=1) + randn(*(tuple(batch_shape) + (f,))) * 0.1
    mask = np.rint(rng.uniform(0.4, 1.0, size=nb * f)).reshape(tuple(batch_shape) + (f,))
    mask1d = np.rint(rng.uniform(0.4, 1.0, size=f))
    if dtype is not None:
        A, y = A.astype(dtype), y.astype(dtype)
        rdtype = np.zeros(1, dtype).real.dtype
        mask, mask1d = mask.astype(rdtype), mask1d.astype(rdtype)
    return A, y, mask, mask1d


def lasso_cases():
    cases = OrderedDict()
    A, y, mask, mask1d = _lasso_data((11,), 5, 10, 0)
    Ac, yc, maskc, mask1dc = _lasso_data((11,), 5, 10, 0, complex_=True)
    Ap, yp, maskp, mask1dp = _lasso_data((11,), 5, 10, 0, positive=True)
    for method in ('ista', 'fista', 'acc_ista'):
        for mname, m, mc, mp in (('nomask', None, None, None), ('mask', mask, maskc, maskp),
                                 ('mask1d', mask1d, mask1dc, mask1dp)):
            cases['mat_%s_%s' % (method, mname)] = dict(
                y=y, A=A, alpha=0.1, tol=1.0e-6, method=method, maxiter=1000, mask=m)
            cases['cmat_%s_%s' % (method, mname)] = dict(
                y=yc, A=Ac, alpha=0.1, tol=1.0e-6, method=method, maxiter=1000, mask=mc)
            cases['pmat_%s_pos_%s' % (method, mname)] = dict(
                y=yp, A=Ap, alpha=0.01, tol=1.0e-6, method=method + '_pos', maxiter=1000, mask=mp)
    # vector and tensor batch shapes
    Av, yv, maskv, _ = _lasso_data((), 5, 10, 0)
    cases['vec_fista_nomask'] = dict(y=yv, A=Av, alpha=0.1, tol=1.0e-6, method='fista', maxiter=1000, mask=None)
    cases['vec_ista_mask'] = dict(y=yv, A=Av, alpha=0.1, tol=1.0e-6, method='ista', maxiter=1000, mask=maskv)
    At, yt, maskt, mask1dt = _lasso_data((12, 11), 5, 10, 0)
    cases['ten_fista_nomask'] = dict(y=yt, A=At, alpha=1.0, tol=1.0e-6, method='fista', maxiter=1000, mask=None)
    cases['ten_fista_mask'] = dict(y=yt, A=At, alpha=0.1, tol=1.0e-6, method='fista', maxiter=1000, mask=maskt)
    cases['ten_ista_mask1d'] = dict(y=yt, A=At, alpha=0.1, tol=1.0e-6, method='ista', maxiter=1000, mask=mask1dt)
    # float32 inputs (reference: tests/test_lasso.py:268-281)
    Af, yf, maskf, _ = _lasso_data((11,), 5, 10, 0, dtype=np.float32)
    cases['f32_fista_nomask'] = dict(y=yf, A=Af, alpha=0.1, tol=1.0e-4, method='fista', maxiter=1000, mask=None)
    cases['f32_ista_mask'] = dict(y=yf, A=Af, alpha=0.1, tol=1.0e-4, method='ista', maxiter=1000, mask=maskf)
    # fixed iteration counts (tol=0: never converges), under- and over-complete, ragged sizes
    Ab, yb, maskb, mask1db = _lasso_data((300,), 24, 40, 1)
    for method in ('ista', 'fista', 'acc_ista', 'fista_pos'):
        cases['fix_%s_nomask' % method] = dict(y=yb, A=Ab, alpha=0.05, tol=0.0, method=method, maxiter=57, mask=None)
        cases['fix_%s_mask' % method] = dict(y=yb, A=Ab, alpha=0.05, tol=0.0, method=method, maxiter=57, mask=maskb)
    cases['fix_fista_mask1d'] = dict(y=yb, A=Ab, alpha=0.05, tol=0.0, method='fista', maxiter=57, mask=mask1db)
    Ao, yo, masko, _ = _lasso_data((129,), 37, 19, 2)          # k > f (ill-conditioned, tests/test_lasso.py:375-420)
    cases['over_fista_nomask'] = dict(y=yo, A=Ao, alpha=0.1, tol=0.0, method='fista', maxiter=40, mask=None)
    cases['over_fista_mask'] = dict(y=yo, A=Ao, alpha=0.1, tol=0.0, method='fista', maxiter=40, mask=masko)
    Aoc, yoc, maskoc, _ = _lasso_data((65,), 9, 33, 4, complex_=True)
    cases['cfix_fista_nomask'] = dict(y=yoc, A=Aoc, alpha=0.05, tol=0.0, method='fista', maxiter=45, mask=None)
    cases['cfix_fista_mask'] = dict(y=yoc, A=Aoc, alpha=0.05, tol=0.0, method='fista', maxiter=45, mask=maskoc)
    cases['cfix_ista_nomask'] = dict(y=yoc, A=Aoc, alpha=0.05, tol=0.0, method='ista', maxiter=45, mask=None)
    return cases


# ------------------------------------------------------------------------------- dictionary learning
def _dl_data(n, f, k, seed, complex_=False):
    """/root/reference/tests/test_dictionary.py:35-43, 80-82."""
    rng = np.random.RandomState(seed)

    def randn(*shape):
        if complex_:
            return rng.randn(*shape) + rng.randn(*shape) * 1.0j
        return rng.randn(*shape)

    Dt = randn(k, f)
    xt = randn(n, k) * rng.uniform(size=n * k).reshape(n, k)
    y = np.dot(xt, Dt) + randn(n, f) * 0.1
    D0 = Dt + randn(k, f) * 0.2
    mask = np.rint(rng.uniform(0.45, 1, size=n * f)).reshape(n, f)
    return y, D0, mask


def dl_cases():
    cases = OrderedDict()
    for cname, cplx in (('f', False), ('c', True)):
        y, D0, mask = _dl_data(101, 5, 3, 0, cplx)
        for lm in ('fista', 'ista', 'acc_ista'):
            cases['%s_%s_conv' % (cname, lm)] = dict(
                y=y, D=D0, alpha=0.1, tol=1.0e-4, minibatch=100, maxiter=1000, lasso_method=lm,
                lasso_iter=1000, lasso_tol=1.0e-5, mask=None, random_seed=0)
            cases['%s_%s_conv_mask' % (cname, lm)] = dict(
                y=y * mask, D=D0, alpha=0.1, tol=1.0e-4, minibatch=100, maxiter=1000, lasso_method=lm,
                lasso_iter=1000, lasso_tol=1.0e-5, mask=mask, random_seed=0)
        # several minibatches per epoch, one tail row dropped each epoch, fixed number of epochs
        cases['%s_fista_mb10' % cname] = dict(
            y=y, D=D0, alpha=0.1, tol=0.0, minibatch=10, maxiter=6, lasso_method='fista',
            lasso_iter=10, lasso_tol=1.0e-5, mask=None, random_seed=3)
        cases['%s_fista_mb10_mask' % cname] = dict(
            y=y * mask, D=D0, alpha=0.1, tol=0.0, minibatch=10, maxiter=6, lasso_method='fista',
            lasso_iter=10, lasso_tol=1.0e-5, mask=mask, random_seed=3)
    y, D0, mask = _dl_data(203, 21, 9, 5, False)
    cases['f_ragged_fista'] = dict(
        y=y, D=D0, alpha=0.05, tol=0.0, minibatch=64, maxiter=4, lasso_method='fista',
        lasso_iter=10, lasso_tol=1.0e-5, mask=None, random_seed=1)
    cases['f_ragged_fista_mask'] = dict(
        y=y * mask, D=D0, alpha=0.05, tol=0.0, minibatch=64, maxiter=4, lasso_method='fista',
        lasso_iter=10, lasso_tol=1.0e-5, mask=mask, random_seed=1)
    y, D0, mask = _dl_data(150, 17, 6, 6, True)
    cases['c_ragged_ista_mask'] = dict(
        y=y * mask, D=D0, alpha=0.05, tol=0.0, minibatch=32, maxiter=4, lasso_method='ista',
        lasso_iter=10, lasso_tol=1.0e-5, mask=mask, random_seed=2)
    return cases
