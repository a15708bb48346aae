"""Host-side behaviour that needs no GPU: the C-ABI library loads and exports every declared symbol, argument
validation raises the reference's exceptions before any device work, and the product path fails loudly (no CPU
fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from decomp_b200 import _lib, dictionary_learning, lasso, nmf
from decomp_b200.utils.exceptions import DimInvalidError, DtypeMismatchError, ShapeMismatchError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAS_GPU = torch.cuda.is_available()


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'decomp_b200.h')).read()
    declared = set(re.findall(r'^(?:const char\*|int|size_t)\s+(decomp_[a-z0-9_]+)\s*\(', header, re.M))
    assert declared, 'no declarations found'
    assert declared == set(_lib.EXPORTS), (sorted(declared - set(_lib.EXPORTS)), sorted(set(_lib.EXPORTS) - declared))
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib.decomp_abi_version.restype = ctypes.c_int
    assert lib.decomp_abi_version() == 2


def test_epilogue_struct_matches_header_layout(tmp_path):
    """The ctypes mirror of decomp_epilogue_t has the size and field offsets gcc gives the C declaration."""
    import shutil
    import subprocess
    if shutil.which('gcc') is None:
        pytest.skip('gcc not available')
    fields = [f[0] for f in _lib.Epilogue._fields_]
    src = tmp_path / 'layout.c'
    body = ''.join('  printf("%s %%zu\\n", offsetof(decomp_epilogue_t, %s));\n' % (f, f) for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "decomp_b200.h"\nint main(void) {\n'
                   '  printf("sizeof %zu\\n", sizeof(decomp_epilogue_t));\n' + body + '  return 0;\n}\n')
    exe = tmp_path / 'layout'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    out = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    assert int(out['sizeof']) == ctypes.sizeof(_lib.Epilogue)
    for f in fields:
        assert int(out[f]) == getattr(_lib.Epilogue, f).offset, f


def test_exceptions_are_value_errors():
    for e in (ShapeMismatchError, DimInvalidError, DtypeMismatchError):
        assert issubclass(e, ValueError)


# ---- validation mirrors /root/reference/tests/test_lasso.py:59-123
def test_lasso_validation_errors():
    rng = np.random.RandomState(0)
    A, y = rng.randn(5, 10), rng.randn(11, 10)
    with pytest.raises(ShapeMismatchError):
        lasso.solve(y, rng.randn(5, 9), 0.1)
    with pytest.raises(ShapeMismatchError):
        lasso.solve(y, A, 0.1, x=rng.randn(11, 4))
    with pytest.raises(ShapeMismatchError):
        lasso.solve(y, A, 0.1, x=rng.randn(10, 5))
    with pytest.raises(ShapeMismatchError):
        lasso.solve(y, A, 0.1, mask=np.ones((11, 9)))
    with pytest.raises(ShapeMismatchError):
        lasso.solve(y, A, 0.1, mask=np.ones(9))
    with pytest.raises(DimInvalidError):
        lasso.solve(y, rng.randn(2, 5, 10), 0.1)
    with pytest.raises(DtypeMismatchError):
        lasso.solve(y, A.astype(np.float32), 0.1)
    with pytest.raises(DtypeMismatchError):
        lasso.solve(y, A + 0j, 0.1)
    with pytest.raises(DtypeMismatchError):
        lasso.solve(y, A, 0.1, mask=np.ones((11, 10), dtype=int))
    with pytest.raises(DtypeMismatchError):
        lasso.solve(y, A, 0.1, mask=np.ones((11, 10), dtype=complex))
    with pytest.raises(AssertionError):
        lasso.solve(y, A, 0.1, mask=-np.ones((11, 10)))
    with pytest.raises(ValueError):
        lasso.solve(y, A, 0.1, method='nonsense')
    with pytest.raises(AssertionError):
        lasso.solve(y + 0j, A + 0j, 0.1, method='ista_pos')
    with pytest.raises(TypeError):
        lasso.solve(y, torch.from_numpy(A), 0.1)


@pytest.mark.parametrize('method', ['cd', 'parallel_cd', 'admm', 'cd_pos'])
def test_lasso_methods_outside_the_hot_path(method):
    rng = np.random.RandomState(0)
    with pytest.raises(NotImplementedError):
        lasso.solve(rng.randn(11, 10), rng.randn(5, 10), 0.1, method=method)


# ---- mirrors decomp/nmf.py:56-68
def test_nmf_validation_errors():
    rng = np.random.RandomState(0)
    y, D = np.abs(rng.randn(20, 7)), np.abs(rng.randn(3, 7))
    with pytest.raises(ShapeMismatchError):
        nmf.solve(y, np.abs(rng.randn(3, 6)))
    with pytest.raises(ShapeMismatchError):
        nmf.solve(y, D, x=np.ones((20, 4)))
    with pytest.raises(ShapeMismatchError):
        nmf.solve(y, D, mask=np.ones((20, 6)))
    with pytest.raises(DtypeMismatchError):
        nmf.solve(y, D.astype(np.float32))
    with pytest.raises(DtypeMismatchError):
        nmf.solve(y + 0j, D + 0j)                      # NMF is real-only (nmf.py:57)
    with pytest.raises(DtypeMismatchError):
        nmf.solve(y, D, mask=np.ones((20, 7), dtype=np.float32))
    with pytest.raises(DimInvalidError):
        nmf.solve(np.stack([y, y]), D, x=np.ones((20, 3)))
    with pytest.raises(AssertionError):
        nmf.solve(y, -D)
    with pytest.raises(AssertionError):
        nmf.solve(y, D, x=-np.ones((20, 3)))
    with pytest.raises(AssertionError):
        nmf.solve(-y, D, likelihood='kl')
    with pytest.raises(NotImplementedError):
        nmf.solve(y, D, method='mu', minibatch=5)                 # 'mu' is a batch method (nmf.py:113)
    with pytest.raises(ValueError):
        nmf.solve(y, D, method='svrmu', minibatch=50)             # minibatch > n (utils/data.py:79-82)
    with pytest.raises(TypeError):
        nmf.solve(y, D, method='svrmu', minibatch=5, forget_rate=0.5)
    with pytest.raises(NotImplementedError):
        nmf.solve(y, D, method='als')
    with pytest.raises(NotImplementedError):
        nmf.solve(y, D, likelihood='huber')
    with pytest.raises(TypeError):
        nmf.solve(y, D, bogus=1)


# ---- mirrors decomp/dictionary_learning.py:65-74,110 and utils/data.py:79-82
def test_dictionary_learning_validation_errors():
    rng = np.random.RandomState(0)
    y, D = rng.randn(20, 7), rng.randn(3, 7)
    with pytest.raises(NotImplementedError):
        dictionary_learning.solve(y, D, 0.1)                                  # minibatch required
    with pytest.raises(ValueError):
        dictionary_learning.solve(y, D, 0.1, minibatch=21, lasso_method='fista')
    with pytest.raises(ShapeMismatchError):
        dictionary_learning.solve(y, rng.randn(3, 6), 0.1, minibatch=10, lasso_method='fista')
    with pytest.raises(DtypeMismatchError):
        dictionary_learning.solve(y, D + 0j, 0.1, minibatch=10, lasso_method='fista')
    with pytest.raises(DtypeMismatchError):
        dictionary_learning.solve(y, D, 0.1, minibatch=10, lasso_method='fista', mask=np.ones((20, 7), dtype=int))
    with pytest.raises(NotImplementedError):
        dictionary_learning.solve(y, D, 0.1, minibatch=10, method='parallel_cd', lasso_method='fista')
    with pytest.raises(NotImplementedError):
        dictionary_learning.solve(y, D, 0.1, minibatch=10)                    # default lasso_method='cd'


@pytest.mark.skipif(HAS_GPU, reason='checks the no-GPU behaviour')
def test_product_path_fails_loudly_without_cuda():
    rng = np.random.RandomState(0)
    with pytest.raises(_lib.DecompError):
        lasso.solve(rng.randn(11, 10), rng.randn(5, 10), 0.1)
    with pytest.raises(_lib.DecompError):
        nmf.solve(np.abs(rng.randn(20, 7)), np.abs(rng.randn(3, 7)))
    with pytest.raises(_lib.DecompError):
        dictionary_learning.solve(rng.randn(20, 7), rng.randn(3, 7), 0.1, minibatch=10, lasso_method='fista')


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, 'decomp_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert 'oracle' not in src.replace('the numpy oracle', ''), os.path.join(dirpath, f)


def test_nnls_wrapper_keeps_the_reference_quirk():
    from decomp_b200 import nnls
    rng = np.random.RandomState(0)
    with pytest.raises(ValueError):
        nnls.solve(rng.randn(11, 10), rng.randn(5, 10), 0.1)          # default 'ista_pos' -> 'ista_pos_pos'
