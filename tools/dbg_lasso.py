import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import golden_cases as gc
from decomp_b200 import lasso
for name, case in gc.lasso_cases().items():
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'lasso_' + name + '.npz'))
    try:
        it, x = lasso.solve(case['y'], case['A'], case['alpha'], tol=case['tol'], method=case['method'],
                            maxiter=case['maxiter'], mask=case['mask'])
        err = np.max(np.abs(x - g['x'])) / max(np.max(np.abs(g['x'])), 1e-300)
        print('%-32s it %4d gold %4d relerr %.3g' % (name, it, int(g['it']), err))
    except Exception as e:
        print(name, 'EXC', repr(e)[:200])
