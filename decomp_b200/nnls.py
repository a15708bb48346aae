"""Non-negative least squares: the reference's thin wrapper over the Lasso solvers (decomp/nnls.py:4-7).

The reference appends ``'_pos'`` to ``method`` and calls ``lasso.solve``; its default ``method='ista_pos'``
therefore becomes ``'ista_pos_pos'`` and raises ``ValueError``.  The same
behaviour is kept, so callers pass the base rule: ``nnls.solve(y, A, alpha, method='fista')``.
"""
from . import lasso


def solve(y, A, alpha, x=None, tol=1.0e-3, method='ista_pos', maxiter=1000, mask=None, **kwargs):
    return lasso.solve(y, A, alpha, x=x, tol=tol, method=method + '_pos', maxiter=maxiter, mask=mask, **kwargs)
