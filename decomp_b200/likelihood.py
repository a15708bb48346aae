"""The reference's finer operator seam for NMF: a ``Likelihood`` whose gradients run on the
GPU, to be passed as ``likelihood=`` to the UNMODIFIED reference ``decomp.nmf.solve`` / ``batch_mu.solve``
(reference: decomp/nmf_methods/grads.py:17-93 the ABC, :108-125 the Gaussian gradients).

    import decomp                                   # the reference
    from decomp.nmf_methods import grads
    from decomp_b200.likelihood import gaussian
    it, D, x = decomp.nmf.solve(y, D0, likelihood=gaussian(grads.Likelihood)(), mask=mask)

It is a parity rig, not the fast path: the reference's driver keeps ``x`` and ``D`` on the host and calls the
likelihood twice per sweep, so every call uploads ``x`` and ``D`` and downloads a gradient pair; ``y`` and ``mask``
(the same objects on every call) are uploaded once and cached.  The whole-solve entry point ``decomp_b200.nmf.solve``
is the product.  The package does not import the reference: the base class is handed in.
"""
import weakref

from . import ops
from ._device import empty2d, np_dtype, require_cuda, to_device2d, to_host


class _DeviceCache(object):
    """Device copies of host arrays that are passed again and again (y, mask), keyed by object identity."""

    def __init__(self):
        self.items = {}

    def get(self, a, device, build):
        key = id(a)
        hit = self.items.get(key)
        if hit is not None and hit[0]() is a:
            return hit[1]
        t = build()
        try:
            self.items[key] = (weakref.ref(a, lambda _r, k=key: self.items.pop(k, None)), t)
        except TypeError:                 # not weak-referenceable: no caching
            pass
        return t


def gaussian(base):
    """A subclass of the reference's ``Likelihood`` (``base``) with the Gaussian / l2 gradients of grads.py:108-125
    computed by the FP64 tensor-core GEMMs of this package.  ``update_x`` / ``update_d`` are inherited from ``base``
    (the reference's own ``x * max(pos, 0) / max(neg, 1e-15)``)."""

    class B200Gaussian(base):
        def __init__(self):
            base.__init__(self)
            self._cache = _DeviceCache()

        # ---- helpers
        def _inputs(self, y, x, d, mask):
            dev = require_cuda()
            if mask is None:
                ym = self._cache.get(y, dev, lambda: to_device2d(y, dev, copy=True))
                md = None
            else:
                md = self._cache.get(mask, dev, lambda: to_device2d(mask, dev, copy=True))

                def masked_y():
                    t = to_device2d(y, dev, copy=True)
                    ops.mask_mul(t, md, t)                       # y * mask, once (grads.py:113,123)
                    return t
                ym = self._cache.get(y, dev, masked_y)
            return dev, ym, md, to_device2d(x, dev, copy=True), to_device2d(d, dev, copy=True)

        def _f(self, dev, xd, dd, md):
            """f = x d (* mask): the [n, f] intermediate of the masked gradients."""
            n, f = xd.shape[0], dd.shape[1]
            dt = ops.make_rhs(dd, False, False)                   # d^T as the NT operand of x . d
            F = empty2d(n, f, False, dev)
            kind = ops.EPI_STORE if md is None else ops.EPI_STORE_MASK
            ops.gemm_nt(xd, dt, ops.epilogue(kind, F, mask=md))
            return F

        # ---- the seam (grads.py:108-125)
        def grad_x(self, y, x, d, mask):
            dev, ym, md, xd, dd = self._inputs(y, x, d, mask)
            n, k = xd.shape
            pos, neg = empty2d(n, k, False, dev), empty2d(n, k, False, dev)
            ops.gemm_nt(ym, dd, ops.epilogue(ops.EPI_STORE, pos))             # (y [* mask]) d^T
            if md is None:
                G = empty2d(k, k, False, dev)
                ops.gemm_nt(dd, dd, ops.epilogue(ops.EPI_STORE, G))           # f d^T = x (d d^T)
                ops.gemm_nt(xd, G, ops.epilogue(ops.EPI_STORE, neg))
            else:
                ops.gemm_nt(self._f(dev, xd, dd, md), dd, ops.epilogue(ops.EPI_STORE, neg))
            dt = np_dtype(y)
            return to_host(pos, y, dt), to_host(neg, y, dt)

        def grad_d(self, y, x, d, mask):
            dev, ym, md, xd, dd = self._inputs(y, x, d, mask)
            n, k = xd.shape
            f = dd.shape[1]
            pos, neg = empty2d(k, f, False, dev), empty2d(k, f, False, dev)
            ws = ops.gemm_tn_workspace_for([(k, f, n), (k, k, n)], dev)
            ops.gemm_tn(xd, ym, pos, workspace=ws)                            # x^T (y [* mask])
            if md is None:
                S = empty2d(k, k, False, dev)
                ops.gemm_tn(xd, xd, S, workspace=ws)                          # x^T f = (x^T x) d
                ops.gemm_nt(S, ops.make_rhs(dd, False, False), ops.epilogue(ops.EPI_STORE, neg))
            else:
                ops.gemm_tn(xd, self._f(dev, xd, dd, md), neg, workspace=ws)
            dt = np_dtype(y)
            return to_host(pos, y, dt), to_host(neg, y, dt)

    return B200Gaussian
