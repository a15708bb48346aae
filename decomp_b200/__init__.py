"""decomp_b200: B200-native (sm_100a) implementation of the deComP iterative-decomposition hot path.

Drop-in for the reference's ``decomp.nmf.solve``, ``decomp.lasso.solve`` and
``decomp.dictionary_learning.solve`` (same signatures, return tuples and exceptions); the
arithmetic runs in hand-written CUDA kernels reached through the C ABI of
``include/decomp_b200.h``.  There is no CPU fallback.
"""
__version__ = '0.1.0'

from . import utils  # noqa: F401,E402
from . import lasso, nmf, dictionary_learning, nnls  # noqa: F401,E402
