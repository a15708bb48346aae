"""Argument validation with the reference's error behaviour (decomp/utils/assertion.py).

Works on numpy arrays and torch tensors alike (only ``shape``, ``ndim``/``dim()`` and
``dtype`` are inspected; the sign checks use the array's own comparison operators).
"""
import numpy as np

from .exceptions import DimInvalidError, DtypeMismatchError, ShapeMismatchError


def _kind(a):
    """'f' or 'c' (or numpy's kind letter for anything else)."""
    name = str(a.dtype).replace('torch.', '')
    try:
        return np.dtype(name).kind
    except TypeError:
        return '?'


def _dtype_name(a):
    return str(a.dtype).replace('torch.', '')


def _ndim(a):
    return a.ndim if hasattr(a, 'ndim') else a.dim()


def assert_shapes(x_name, x, y_name, y, axes=None):
    """axes None: identical shapes; int n: x.shape[-n:] == y.shape[:n]; list: equal on those axes.
    Either array may be None, in which case nothing is checked (reference: assertion.py:5-42)."""
    if x is None or y is None:
        return
    xs, ys = tuple(x.shape), tuple(y.shape)
    if axes is None:
        ok = xs == ys
        rule = 'Shapes of {0} and {1} should be identical.'.format(x_name, y_name)
    elif isinstance(axes, int):
        ok = xs[-axes:] == ys[:axes]
        rule = '{0}.shape[-{2}:] == {1}.shape[:{2}] should be satisfied.'.format(x_name, y_name, axes)
    elif isinstance(axes, (list, tuple)):
        try:
            ok = all(xs[a] == ys[a] for a in axes)
        except IndexError:
            ok = False
        rule = '{0}.shape[{2}] == {1}.shape[{2}] should be satisfied.'.format(x_name, y_name, list(axes))
    else:
        raise TypeError('Argument axes is invalid, given ' + str(axes))
    if not ok:
        raise ShapeMismatchError('{0} Given {1}: {2} and {3}: {4}'.format(rule, x_name, xs, y_name, ys))


def assert_ndim(x_name, x, ndim):
    if x is None:
        return
    if _ndim(x) != ndim:
        raise DimInvalidError('Dimension of {0} should be {1} but given {2}'.format(x_name, ndim, _ndim(x)))


def assert_dtypes(dtypes='fc', **arrays):
    """All non-None arrays share one dtype, whose kind is in ``dtypes`` (reference: assertion.py:54-84)."""
    given = [(k, a) for k, a in arrays.items() if a is not None]
    if not given:
        return
    k0, a0 = given[0]
    for k, a in given:
        if _dtype_name(a) != _dtype_name(a0):
            raise DtypeMismatchError('Data type should be all identical, {0}: {1} and {2}: {3}'.format(
                k0, _dtype_name(a0), k, _dtype_name(a)))
        if _kind(a) not in dtypes:
            raise DtypeMismatchError('Data type should be one of {0}, but given {1} for {2}'.format(
                dtypes, _dtype_name(a), k))


def assert_nonnegative(x):
    """AssertionError unless x is real and has no negative entry (reference: assertion.py:95-100)."""
    if x is None:
        return
    assert _kind(x) != 'c'
    assert bool((x >= 0.0).all())
