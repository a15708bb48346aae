"""Measured parity errors (relative max-norm) of every golden / oracle case of the public solve() entry points, written
as JSON: which cases sit where below the 1e-10 bar of BASELINE.json's north_star.  Run on the GPU box:

    python tools/parity_errors.py > gpurun_out/parity_errors.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np  # noqa: E402

import golden_cases as gc  # noqa: E402
from decomp_b200 import dictionary_learning, lasso, nmf  # noqa: E402
from oracle import decomp_oracle as orc  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


def rel(a, b):
    a, b = np.asarray(a, dtype=np.complex128), np.asarray(b, dtype=np.complex128)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


out = {'nmf': {}, 'nmf_minibatch': {}, 'lasso': {}, 'dl': {}, 'dl_oracle': {}, 'dl_growth': {}}
for name, case in gc.nmf_cases().items():
    g = np.load(os.path.join(GOLD, 'nmf_%s.npz' % name))
    it, D, x = nmf.solve(case['y'], case['D'].copy(), tol=case['tol'], maxiter=case['maxiter'],
                         likelihood=case['likelihood'], mask=case['mask'])
    out['nmf'][name] = dict(it=it, it_ref=int(g['it']), err_D=rel(D, g['D']), err_x=rel(x, g['x']))
for name, case in gc.nmf_minibatch_cases().items():
    g = np.load(os.path.join(GOLD, 'nmfmb_%s.npz' % name))
    it, D, x = nmf.solve(case['y'], case['D'].copy(), tol=case['tol'], minibatch=case['minibatch'],
                         maxiter=case['maxiter'], method=case['method'], likelihood=case['likelihood'],
                         mask=case['mask'], random_seed=case['random_seed'])
    out['nmf_minibatch'][name] = dict(it=it, it_ref=int(g['it']), err_D=rel(D, g['D']), err_x=rel(x, g['x']))
for name, case in gc.lasso_cases().items():
    g = np.load(os.path.join(GOLD, 'lasso_%s.npz' % name))
    it, x = lasso.solve(case['y'], case['A'], case['alpha'], tol=case['tol'], method=case['method'],
                        maxiter=case['maxiter'], mask=case['mask'])
    out['lasso'][name] = dict(it=it, it_ref=int(g['it']), err_x=rel(x, g['x']))
for name, case in gc.dl_cases().items():
    g = np.load(os.path.join(GOLD, 'dl_%s.npz' % name))
    kw = {k: v for k, v in case.items() if k not in ('y', 'D', 'alpha')}
    it, D, x = dictionary_learning.solve(case['y'], case['D'].copy(), case['alpha'], **kw)
    out['dl'][name] = dict(it=it, it_ref=int(g['it']), err_D=rel(D, g['D']), err_x=rel(x, g['x']))
for cplx in (False, True):
    for masked in (False, True):
        y, D0, mask = gc._dl_data(1030, 75, 40, 9, cplx)
        m = mask if masked else None
        yy = y * mask if masked else y
        key = '%s_%s' % ('c128' if cplx else 'f64', 'mask' if masked else 'nomask')
        growth = []
        for maxiter in (2, 3, 5):
            kw = dict(tol=0.0, minibatch=256, maxiter=maxiter, lasso_method='fista', lasso_iter=10, lasso_tol=1.0e-5,
                      mask=m, random_seed=4)
            it0, D_ref, x_ref = orc.dictionary_learning(yy, D0.copy(), 0.05, **kw)
            it, D, x = dictionary_learning.solve(yy, D0.copy(), 0.05, **kw)
            growth.append(dict(epochs=maxiter - 1, err_D=rel(D, D_ref), err_x=rel(x, x_ref)))
        out['dl_growth'][key] = growth
worst = {}
for fam, cases in out.items():
    if fam == 'dl_growth':
        continue
    errs = [max(v.get('err_D', 0.0), v.get('err_x', 0.0)) for v in cases.values()]
    if errs:
        worst[fam] = max(errs)
out['worst'] = worst
print(json.dumps(out, indent=1))
