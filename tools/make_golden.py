#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (numpy path).

Runs only in the build container, where /root/reference is mounted.  The reference's
``decomp`` package hard-imports ``chainer`` (decomp/utils/cp_compat.py:1,
decomp/template_matching.py:2) although the hot path never uses it, so an empty stub
package is created in a temporary directory and put on ``sys.path``; no reference file
is copied or modified.  Inputs are re-creatable from the seeds recorded in each file
(``tests/golden_cases.py`` holds the generators shared with the tests), outputs are
what the reference returned.

    python tools/make_golden.py            # rewrites tests/golden/
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
REFERENCE = os.environ.get('DECOMP_REFERENCE', '/root/reference')


def import_reference():
    stub = tempfile.mkdtemp(prefix='chainer_stub_')
    os.makedirs(os.path.join(stub, 'chainer', 'utils'))
    files = {
        'chainer/__init__.py': 'from . import cuda\n',
        'chainer/cuda.py': '',
        'chainer/utils/__init__.py': 'from . import conv, conv_nd\n',
        'chainer/utils/conv.py': '',
        'chainer/utils/conv_nd.py': '',
    }
    for name, body in files.items():
        with open(os.path.join(stub, name), 'w') as f:
            f.write(body)
    sys.path.insert(0, stub)
    sys.path.insert(0, REFERENCE)
    import decomp  # noqa: E402
    return decomp


def main():
    import golden_cases as gc
    ref = import_reference()
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(out_dir, exist_ok=True)

    # ---- literal known answers of the reference's own tests (tests/test_lasso.py:15-56)
    z = np.array([[0.1, -2.0, 1.4], [1.1, 3.0, -1.4]])
    zc = np.array([0.1, -2.0, 1.4])
    np.savez(os.path.join(out_dir, 'soft_threshold.npz'),
             z=z, real=ref.lasso.soft_threshold_float(z, 1.0, np),
             zc=zc + 1j * zc, cplx=ref.lasso.soft_threshold_complex(zc + 1j * zc, 1.0, np),
             pos=ref.lasso.soft_threshold_positive(z, 1.0, np))

    # ---- NMF
    for name, case in gc.nmf_cases().items():
        it, D, x = ref.nmf.solve(case['y'], case['D'].copy(), x=None, tol=case['tol'],
                                 maxiter=case['maxiter'], method='mu',
                                 likelihood=case['likelihood'], mask=case['mask'])
        np.savez(os.path.join(out_dir, 'nmf_%s.npz' % name), it=it, D=D, x=x)
        print('nmf', name, 'it', it, 'sumD', D.sum(), 'sumx', x.sum())

    # ---- minibatch NMF drivers
    for name, case in gc.nmf_minibatch_cases().items():
        it, D, x = ref.nmf.solve(case['y'], case['D'].copy(), x=None, tol=case['tol'], minibatch=case['minibatch'],
                                 maxiter=case['maxiter'], method=case['method'], likelihood=case['likelihood'],
                                 mask=case['mask'], random_seed=case['random_seed'])
        np.savez(os.path.join(out_dir, 'nmfmb_%s.npz' % name), it=it, D=D, x=x)
        print('nmfmb', name, 'it', it, 'sumD', D.sum(), 'sumx', x.sum())

    # ---- Lasso
    for name, case in gc.lasso_cases().items():
        it, x = ref.lasso.solve(case['y'], case['A'], alpha=case['alpha'], tol=case['tol'],
                                method=case['method'], maxiter=case['maxiter'], mask=case['mask'])
        np.savez(os.path.join(out_dir, 'lasso_%s.npz' % name), it=it, x=x)
        print('lasso', name, 'it', it, 'sum|x|', np.abs(x).sum())

    # ---- dictionary learning
    for name, case in gc.dl_cases().items():
        it, D, x = ref.dictionary_learning.solve(
            case['y'], case['D'].copy(), case['alpha'], x=None, tol=case['tol'],
            minibatch=case['minibatch'], maxiter=case['maxiter'],
            lasso_method=case['lasso_method'], lasso_iter=case['lasso_iter'],
            lasso_tol=case['lasso_tol'], mask=case['mask'], random_seed=case['random_seed'])
        np.savez(os.path.join(out_dir, 'dl_%s.npz' % name), it=it, D=D, x=x)
        print('dl', name, 'it', it, 'sum|D|', np.abs(D).sum())


if __name__ == '__main__':
    main()
