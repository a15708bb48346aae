"""The oracle (oracle/decomp_oracle.py) against outputs of the unmodified reference.

tests/golden/*.npz were written by tools/make_golden.py, which imports the reference
from /root/reference and runs it on the seeded inputs of tests/golden_cases.py.  The
oracle must reproduce iteration counts exactly and factors to rounding.
"""
import os

import numpy as np
import pytest

import golden_cases as gc
from oracle import decomp_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
RTOL = 1.0e-12


def load(name):
    return np.load(os.path.join(GOLD, name + '.npz'))


def close(a, b, rtol=RTOL):
    scale = max(np.max(np.abs(b)), 1.0e-300)
    return np.max(np.abs(a - b)) <= rtol * scale


def test_soft_threshold_known_answers():
    g = load('soft_threshold')
    # literal vectors of /root/reference/tests/test_lasso.py:22-33
    assert np.allclose(orc.shrink_real(np.array([0.1, -2.0, 1.4]), 1.0), [0.0, -1.0, 0.4])
    assert np.allclose(orc.shrink_real(np.array([[0.1, -2.0, 1.4], [1.1, 3.0, -1.4]]), 1.0),
                       [[0.0, -1.0, 0.4], [0.1, 2.0, -0.4]])
    assert np.array_equal(orc.shrink_real(g['z'], 1.0), g['real'])
    assert np.array_equal(orc.shrink_complex(g['zc'], 1.0), g['cplx'])
    assert np.array_equal(orc.shrink_positive(g['z'], 1.0), g['pos'])
    # rotational consistency (/root/reference/tests/test_lasso.py:35-56)
    x = np.array([0.1, -2.0, 1.4])
    assert np.allclose(orc.shrink_complex(x * 1.0j, 1.0), orc.shrink_complex(x + 0j, 1.0) * 1.0j)
    assert np.allclose(orc.shrink_complex(x + x * 1.0j, 1.0),
                       orc.shrink_complex(x * np.sqrt(2.0) + 0j, 1.0) * (np.sqrt(0.5) + np.sqrt(0.5) * 1.0j))


@pytest.mark.parametrize('name', list(gc.nmf_cases().keys()))
def test_nmf(name):
    case = gc.nmf_cases()[name]
    g = load('nmf_' + name)
    it, D, x = orc.nmf_mu(case['y'], case['D'].copy(), tol=case['tol'], maxiter=case['maxiter'],
                          likelihood=case['likelihood'], mask=case['mask'])
    assert it == int(g['it'])
    assert close(D, g['D']) and close(x, g['x'])


@pytest.mark.parametrize('name', list(gc.nmf_minibatch_cases().keys()))
def test_nmf_minibatch(name):
    case = gc.nmf_minibatch_cases()[name]
    g = load('nmfmb_' + name)
    kw = {k: v for k, v in case.items() if k not in ('y', 'D')}
    it, D, x = orc.nmf_minibatch(case['y'], case['D'].copy(), **kw)
    assert it == int(g['it'])
    assert close(D, g['D']) and close(x, g['x'])


@pytest.mark.parametrize('name', list(gc.lasso_cases().keys()))
def test_lasso(name):
    case = gc.lasso_cases()[name]
    g = load('lasso_' + name)
    it, x = orc.lasso(case['y'], case['A'], case['alpha'], tol=case['tol'], method=case['method'],
                      maxiter=case['maxiter'], mask=case['mask'])
    assert it == int(g['it'])
    assert x.dtype == g['x'].dtype and x.shape == g['x'].shape
    assert close(x, g['x'], 1.0e-6 if x.dtype in (np.float32, np.complex64) else RTOL)


@pytest.mark.parametrize('name', list(gc.dl_cases().keys()))
def test_dictionary_learning(name):
    case = gc.dl_cases()[name]
    g = load('dl_' + name)
    kw = {k: v for k, v in case.items() if k not in ('y', 'D', 'alpha')}
    it, D, x = orc.dictionary_learning(case['y'], case['D'].copy(), case['alpha'], **kw)
    assert it == int(g['it'])
    assert close(D, g['D']) and close(x, g['x'])


def test_survey_anchors():
    """Known-answer anchors recorded in SURVEY.md section 8(c) from the reference."""
    g = load('nmf_test_l2')
    case = gc.nmf_cases()['test_l2']
    assert int(g['it']) == 1484
    assert abs(orc.nmf_objective(case['y'], g['x'], g['D']) - 8.435593463594738) < 1e-9
    assert abs(g['D'].sum() - 7.5305383921848685) < 1e-12
    assert int(load('lasso_mat_ista_nomask')['it']) == 90
    assert int(load('lasso_mat_fista_mask')['it']) == 240
    assert int(load('dl_f_acc_ista_conv')['it']) == 53
    assert int(load('dl_f_acc_ista_conv_mask')['it']) == 125
