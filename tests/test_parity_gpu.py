"""Parity of the public solve() entry points (CUDA path, through the C ABI) against

  * the golden outputs of the unmodified reference (tests/golden/, made by tools/make_golden.py), and
  * the numpy oracle on larger seeded inputs the oracle finishes in seconds.

Tolerance: 1e-10 relative (max-norm) on the returned factors for the FP64 path, as BASELINE.json's
north_star states; iteration counts must be identical.  float32 inputs are widened to FP64 on the
device, so they are compared with the reference's float32 result at float32 accuracy.
"""
import os

import numpy as np
import pytest

import golden_cases as gc

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
RTOL = 1.0e-10


def load(name):
    return np.load(os.path.join(GOLD, name + '.npz'))


def rel_err(a, b):
    scale = max(np.max(np.abs(b)), 1.0e-300)
    return np.max(np.abs(a - b)) / scale


def assert_close(a, b, rtol=RTOL, what=''):
    assert a.shape == b.shape, '%s shape %s vs %s' % (what, a.shape, b.shape)
    assert a.dtype == b.dtype, '%s dtype %s vs %s' % (what, a.dtype, b.dtype)
    err = rel_err(a, b)
    assert err <= rtol, '%s relative error %.3g > %.1g' % (what, err, rtol)


# ----------------------------------------------------------------------------------- golden: NMF
@pytest.mark.parametrize('name', list(gc.nmf_cases().keys()))
def test_nmf_golden(name):
    from decomp_b200 import nmf
    case = gc.nmf_cases()[name]
    g = load('nmf_' + name)
    it, D, x = nmf.solve(case['y'], case['D'].copy(), tol=case['tol'], maxiter=case['maxiter'],
                         likelihood=case['likelihood'], mask=case['mask'])
    assert isinstance(it, int) and it == int(g['it'])
    assert_close(D, g['D'], what='D')
    assert_close(x, g['x'], what='x')


@pytest.mark.parametrize('name', list(gc.nmf_minibatch_cases().keys()))
def test_nmf_minibatch_golden(name):
    """Stochastic MU drivers (serizel.py, kasai.py) against the unmodified reference."""
    from decomp_b200 import nmf
    case = gc.nmf_minibatch_cases()[name]
    g = load('nmfmb_' + name)
    it, D, x = nmf.solve(case['y'], case['D'].copy(), tol=case['tol'], minibatch=case['minibatch'],
                         maxiter=case['maxiter'], method=case['method'], likelihood=case['likelihood'],
                         mask=case['mask'], random_seed=case['random_seed'])
    assert it == int(g['it'])
    assert_close(D, g['D'], what='D')
    assert_close(x, g['x'], what='x')


# ----------------------------------------------------------------------------------- golden: Lasso
@pytest.mark.parametrize('name', list(gc.lasso_cases().keys()))
def test_lasso_golden(name):
    from decomp_b200 import lasso
    case = gc.lasso_cases()[name]
    g = load('lasso_' + name)
    it, x = lasso.solve(case['y'], case['A'], case['alpha'], tol=case['tol'], method=case['method'],
                        maxiter=case['maxiter'], mask=case['mask'])
    single = x.dtype in (np.float32, np.complex64)
    if not single:
        assert it == int(g['it'])
    if single:
        # the reference's output dtype for float32 inputs is an accident of numpy scalar promotion (float64 for
        # fista, float32 for ista); we return the input dtype and compare values at float32 accuracy
        assert x.dtype == case['y'].dtype
        assert rel_err(x.astype(np.float64), g['x'].astype(np.float64)) <= 2.0e-4
    else:
        assert_close(x, g['x'], what='x')


# ----------------------------------------------------------------------------------- golden: dictionary learning
@pytest.mark.parametrize('name', list(gc.dl_cases().keys()))
def test_dictionary_learning_golden(name):
    from decomp_b200 import dictionary_learning
    case = gc.dl_cases()[name]
    g = load('dl_' + name)
    kw = {k: v for k, v in case.items() if k not in ('y', 'D', 'alpha')}
    it, D, x = dictionary_learning.solve(case['y'], case['D'].copy(), case['alpha'], **kw)
    assert it == int(g['it'])
    assert_close(D, g['D'], what='D')
    assert_close(x, g['x'], what='x')


# ----------------------------------------------------------------------------------- oracle on larger inputs
@pytest.mark.parametrize('masked', [False, True])
def test_nmf_vs_oracle_multi_tile(masked):
    """Several row tiles, ragged edges, k spanning more than one N tile."""
    from decomp_b200 import nmf
    from oracle import decomp_oracle as orc
    y, D0, mask = gc._nmf_data(3001, 517, 70, 11)
    m = mask if masked else None
    it0, D_ref, x_ref = orc.nmf_mu(y, D0.copy(), tol=0.0, maxiter=21, mask=m)
    it, D, x = nmf.solve(y, D0.copy(), tol=0.0, maxiter=21, mask=m)
    assert it == it0 == 21
    assert_close(D, D_ref, what='D')
    assert_close(x, x_ref, what='x')
    obj, obj_ref = orc.nmf_objective(y, x, D, m), orc.nmf_objective(y, x_ref, D_ref, m)
    assert abs(obj - obj_ref) <= RTOL * abs(obj_ref)


@pytest.mark.parametrize('method', ['ista', 'fista', 'fista_pos'])
@pytest.mark.parametrize('mask_kind', ['nomask', 'mask'])
def test_lasso_vs_oracle_multi_tile(method, mask_kind):
    from decomp_b200 import lasso
    from oracle import decomp_oracle as orc
    A, y, mask, _ = gc._lasso_data((2500,), 150, 333, 7, positive=method.endswith('_pos'))
    m = mask if mask_kind == 'mask' else None
    it0, x_ref = orc.lasso(y, A, 0.05, tol=0.0, method=method, maxiter=30, mask=m)
    it, x = lasso.solve(y, A, 0.05, tol=0.0, method=method, maxiter=30, mask=m)
    assert it == it0 == 29
    assert_close(x, x_ref, what='x')
    obj, obj_ref = orc.lasso_objective(y, A, x, 0.05, m), orc.lasso_objective(y, A, x_ref, 0.05, m)
    assert abs(obj - obj_ref) <= RTOL * abs(obj_ref)


@pytest.mark.parametrize('method', ['ista', 'fista', 'fista_pos'])
@pytest.mark.parametrize('k', [32, 64, 128])
def test_lasso_masked_fused_b2b_vs_oracle(method, k):
    """Per-problem mask with a code width the fused back-to-back kernel covers (one launch per iteration)."""
    from decomp_b200 import lasso
    from oracle import decomp_oracle as orc
    A, y, mask, _ = gc._lasso_data((1300,), k, 333, 7 + k, positive=method.endswith('_pos'))
    it0, x_ref = orc.lasso(y, A, 0.05, tol=1e-7, method=method, maxiter=120, mask=mask)
    it, x = lasso.solve(y, A, 0.05, tol=1e-7, method=method, maxiter=120, mask=mask)
    assert it == it0
    assert_close(x, x_ref, what='x')
    obj, obj_ref = orc.lasso_objective(y, A, x, 0.05, mask), orc.lasso_objective(y, A, x_ref, 0.05, mask)
    assert abs(obj - obj_ref) <= RTOL * abs(obj_ref)


def test_lasso_masked_fused_b2b_complex_vs_oracle():
    from decomp_b200 import lasso
    from oracle import decomp_oracle as orc
    A, y, mask, _ = gc._lasso_data((700,), 32, 130, 8, complex_=True)      # 2k = 64 real columns
    it0, x_ref = orc.lasso(y, A, 0.05, tol=0.0, method='fista', maxiter=25, mask=mask)
    it, x = lasso.solve(y, A, 0.05, tol=0.0, method='fista', maxiter=25, mask=mask)
    assert it == it0
    assert_close(x, x_ref, what='x')


@pytest.mark.parametrize('k', [32, 128])
def test_nmf_masked_fused_b2b_vs_oracle(k):
    from decomp_b200 import nmf
    from oracle import decomp_oracle as orc
    y, D0, mask = gc._nmf_data(2001, 300, k, 13)
    it0, D_ref, x_ref = orc.nmf_mu(y, D0.copy(), tol=0.0, maxiter=16, mask=mask)
    it, D, x = nmf.solve(y, D0.copy(), tol=0.0, maxiter=16, mask=mask)
    assert it == it0 == 16
    assert_close(D, D_ref, what='D')
    assert_close(x, x_ref, what='x')


def test_lasso_complex_vs_oracle_multi_tile():
    from decomp_b200 import lasso
    from oracle import decomp_oracle as orc
    A, y, mask, _ = gc._lasso_data((700,), 90, 130, 8, complex_=True)
    for m in (None, mask):
        it0, x_ref = orc.lasso(y, A, 0.05, tol=0.0, method='fista', maxiter=25, mask=m)
        it, x = lasso.solve(y, A, 0.05, tol=0.0, method='fista', maxiter=25, mask=m)
        assert it == it0
        assert_close(x, x_ref, what='x')


@pytest.mark.parametrize('cplx', [False, True])
@pytest.mark.parametrize('masked', [False, True])
def test_dictionary_learning_vs_oracle(cplx, masked):
    from decomp_b200 import dictionary_learning
    from oracle import decomp_oracle as orc
    y, D0, mask = gc._dl_data(1030, 75, 40, 9, cplx)
    m = mask if masked else None
    yy = y * mask if masked else y
    kw = dict(tol=0.0, minibatch=256, maxiter=3, lasso_method='fista', lasso_iter=10, lasso_tol=1.0e-5, mask=m,
              random_seed=4)
    it0, D_ref, x_ref = orc.dictionary_learning(yy, D0.copy(), 0.05, **kw)
    it, D, x = dictionary_learning.solve(yy, D0.copy(), 0.05, **kw)
    assert it == it0
    assert_close(D, D_ref, what='D')
    assert_close(x, x_ref, what='x')


# ----------------------------------------------------------------------------------- API behaviour
def test_torch_inputs_stay_on_device():
    import torch
    from decomp_b200 import lasso, nmf
    case = gc.lasso_cases()['mat_fista_nomask']
    g = load('lasso_mat_fista_nomask')
    it, x = lasso.solve(torch.from_numpy(case['y']).cuda(), torch.from_numpy(case['A']).cuda(), case['alpha'],
                        tol=case['tol'], method='fista', maxiter=case['maxiter'])
    assert isinstance(x, torch.Tensor) and x.is_cuda and it == int(g['it'])
    assert rel_err(x.cpu().numpy(), g['x']) <= RTOL
    case = gc.nmf_cases()['ragged_l2']
    g = load('nmf_ragged_l2')
    it, D, x = nmf.solve(torch.from_numpy(case['y']).cuda(), torch.from_numpy(case['D']).cuda(), tol=0.0,
                         maxiter=case['maxiter'])
    assert D.is_cuda and x.is_cuda and it == int(g['it'])
    assert rel_err(D.cpu().numpy(), g['D']) <= RTOL


def test_inputs_are_not_mutated():
    from decomp_b200 import nmf
    case = gc.nmf_cases()['ragged_l2_mask']
    y, D, mask = case['y'].copy(), case['D'].copy(), case['mask'].copy()
    x0 = np.ones((y.shape[0], D.shape[0]))
    nmf.solve(y, D, x=x0, tol=0.0, maxiter=5, mask=mask)
    assert np.array_equal(y, case['y']) and np.array_equal(D, case['D']) and np.array_equal(mask, case['mask'])
    assert np.array_equal(x0, np.ones_like(x0))


def test_nnls_wrapper():
    """decomp/nnls.py:4-7 appends '_pos'; equals the golden of lasso.solve(method='fista_pos')."""
    from decomp_b200 import nnls
    case = gc.lasso_cases()['pmat_fista_pos_nomask']
    g = load('lasso_pmat_fista_pos_nomask')
    it, x = nnls.solve(case['y'], case['A'], case['alpha'], tol=case['tol'], method='fista', maxiter=case['maxiter'])
    assert it == int(g['it'])
    assert_close(x, g['x'], what='x')


@pytest.mark.parametrize('masked', [False, True])
def test_nmf_streamed_from_host_matches_resident(masked):
    """Out-of-core mode (host-resident y streamed in row blocks) against the oracle."""
    from decomp_b200 import nmf
    from oracle import decomp_oracle as orc
    y, D0, mask = gc._nmf_data(1003, 130, 24, 12)
    m = mask if masked else None
    it0, D_ref, x_ref = orc.nmf_mu(y, D0.copy(), tol=1e-4, maxiter=300, mask=m)
    for block in (128, 400, 5000):
        it, D, x = nmf.solve(y, D0.copy(), tol=1e-4, maxiter=300, mask=m, host_block_rows=block)
        assert it == it0
        assert_close(D, D_ref, what='D')
        assert_close(x, x_ref, what='x')


@pytest.mark.parametrize('masked', [False, True])
def test_nmf_small_launch_graph_and_eager_paths_agree(masked, monkeypatch):
    """Small problems run whole solves in one cooperative launch; launch-bound ones replay a CUDA graph; the rest
    enqueue sweep by sweep.  All three give the oracle's iteration count and factors."""
    from decomp_b200 import nmf
    from oracle import decomp_oracle as orc
    y, D0, mask = gc._nmf_data(1000, 200, 20, 0, 'l2', reference_order=False)       # BASELINE configs[0]
    m = mask if masked else None
    for tol, maxiter in ((0.0, 101), (1.0e-4, 2000)):
        it0, D_ref, x_ref = orc.nmf_mu(y, D0.copy(), tol=tol, maxiter=maxiter, mask=m)
        for path in ('small', 'graph', 'eager'):
            monkeypatch.setattr(nmf, 'USE_SMALL', path == 'small')
            monkeypatch.setattr(nmf, 'GRAPH_MIN_SWEEPS', 12 if path != 'eager' else 10 ** 9)
            it, D, x = nmf.solve(y, D0.copy(), tol=tol, maxiter=maxiter, mask=m)
            assert it == it0, (path, tol, it, it0)
            assert_close(D, D_ref, what='D ' + path)
            assert_close(x, x_ref, what='x ' + path)


def test_nmf_small_launch_odd_sizes():
    """Rows that do not fill the last CTA, more CTAs than rows, a single atom."""
    from decomp_b200 import nmf
    from oracle import decomp_oracle as orc
    for n, f, k, seed in ((257, 67, 7, 3), (40, 33, 1, 5), (2000, 31, 32, 6)):
        y, D0, mask = gc._nmf_data(n, f, k, seed)
        for m in (None, mask):
            it0, D_ref, x_ref = orc.nmf_mu(y, D0.copy(), tol=0.0, maxiter=31, mask=m)
            it, D, x = nmf.solve(y, D0.copy(), tol=0.0, maxiter=31, mask=m)
            assert it == it0
            assert_close(D, D_ref, what='D')
            assert_close(x, x_ref, what='x')


def test_dictionary_learning_masked_complex_multi_chunk_statistics():
    """BASELINE configs[3] in small: complex128, masked, and enough atoms that the Hermitian half of the [k, f, k]
    statistic is accumulated in several packed GEMMs (4656 (a, b >= a) pairs: >= 3 chunks of the real
    decomp_b200.dictionary_learning._pair_cols width), through dictionary_learning.solve itself."""
    import torch
    from decomp_b200 import dictionary_learning
    from oracle import decomp_oracle as orc
    n, f, k, mb = 1100, 256, 96, 512
    y, D0, mask = gc._dl_data(n, f, k, 21, True)
    pairs = k * (k + 1) // 2
    cap = dictionary_learning._pair_cols(f, 2, torch.device('cuda', 0))
    assert -(-pairs // cap) >= 3, (pairs, cap)
    kw = dict(tol=0.0, minibatch=mb, maxiter=2, lasso_method='fista', lasso_iter=10, lasso_tol=1.0e-5, mask=mask,
              random_seed=2)
    it0, D_ref, x_ref = orc.dictionary_learning(y * mask, D0.copy(), 0.05, **kw)
    it, D, x = dictionary_learning.solve(y * mask, D0.copy(), 0.05, **kw)
    assert it == it0
    assert_close(D, D_ref, what='D')
    assert_close(x, x_ref, what='x')


@pytest.mark.parametrize('masked', [False, True])
def test_likelihood_seam(masked):
    """The reference's operator seam (grads.py:17-93): a Likelihood with GPU gradients, driven by the reference's own
    update rules -- a stand-in for the ABC here, and the unmodified reference driver when baseline/_ref is installed."""
    from decomp_b200.likelihood import gaussian
    from oracle import cpu_arm
    from oracle import decomp_oracle as orc
    y, D0, mask = gc._nmf_data(301, 75, 6, 3)
    m = mask if masked else None

    class Base(object):                      # the two update rules of the reference ABC, restated (grads.py:77-93)
        def __init__(self):
            pass

        def update_x(self, y, x, d, mask):
            p, n = self.grad_x(y, x, d, mask)
            return x * np.maximum(p, 0.0) / np.maximum(n, 1.0e-15)

        def update_d(self, y, x, d, mask):
            p, n = self.grad_d(y, x, d, mask)
            return d * np.maximum(p, 0.0) / np.maximum(n, 1.0e-15)

    lk = gaussian(Base)()
    x = np.ones((y.shape[0], D0.shape[0]))
    D = D0 / np.sqrt(np.sum(D0 * D0, axis=-1, keepdims=True))
    w = np.ones_like(y) if m is None else m
    f = x.dot(D) * w
    for (got_p, got_n), (ref_p, ref_n) in ((lk.grad_x(y, x, D, m), ((y * w).dot(D.T), f.dot(D.T))),
                                           (lk.grad_d(y, x, D, m), (x.T.dot(y * w), x.T.dot(f)))):
        assert np.max(np.abs(got_p - ref_p)) <= 1e-12 * np.max(np.abs(ref_p))
        assert np.max(np.abs(got_n - ref_n)) <= 1e-12 * np.max(np.abs(ref_n))
    # batch_mu's recursion (batch_mu.py:16-24) through the seam against the oracle
    for _ in range(5):
        x = lk.update_x(y, x, D, m)
        Dn = lk.update_d(y, x, D, m)
        D = Dn / np.sqrt(np.sum(Dn * Dn, axis=-1, keepdims=True))
    it0, D_ref, x_ref = orc.nmf_mu(y, D0.copy(), tol=0.0, maxiter=6, mask=m)
    assert np.max(np.abs(D - D_ref)) <= 1e-10 * np.max(np.abs(D_ref))
    assert np.max(np.abs(x - x_ref)) <= 1e-10 * np.max(np.abs(x_ref))
    # inside the unmodified reference's own driver
    arm = cpu_arm.load()
    if arm.kind == 'reference':
        import decomp
        from decomp.nmf_methods import grads
        it_a, D_a, x_a = decomp.nmf.solve(y, D0.copy(), tol=1e-5, maxiter=300, mask=m)
        it_b, D_b, x_b = decomp.nmf.solve(y, D0.copy(), tol=1e-5, maxiter=300, mask=m,
                                          likelihood=gaussian(grads.Likelihood)())
        assert it_a == it_b
        assert np.max(np.abs(D_a - D_b)) <= 1e-10 * np.max(np.abs(D_a))
        assert np.max(np.abs(x_a - x_b)) <= 1e-10 * np.max(np.abs(x_a))
