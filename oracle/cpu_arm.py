"""CPU arm used by ``bench.py`` (``cpu_baseline`` legs and ``--impl reference``) -- TEST INFRASTRUCTURE.

Picks what is timed on the host cores:

* ``kind == "reference"``: the UNMODIFIED reference package, installed into git-ignored ``baseline/_ref/`` by
  ``tools/install_reference.py`` (run by ``__graft_entry__.build()`` in the build container, where ``/root/reference``
  exists; the directory then travels to the GPU box with the repo snapshot).  Called through its own public API
  (``decomp.lasso.solve`` / ``decomp.nmf.solve`` / ``decomp.dictionary_learning.solve``).
* ``kind == "port"``: ``oracle/decomp_oracle.py`` (the numpy restatement pinned to the reference's outputs) when
  ``baseline/_ref`` is absent.

Nothing in ``decomp_b200`` imports this module.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, 'baseline', '_ref')

_arm = None


class CpuArm(object):
    def __init__(self, kind, lasso, nmf, dictionary_learning, where):
        self.kind, self.lasso, self.nmf, self.dictionary_learning, self.where = kind, lasso, nmf, dictionary_learning, where


def set_blas_threads():
    """All host cores this process may use for BLAS, whatever OMP_NUM_THREADS the launcher exported
    (torch.distributed.run sets it to 1).  Returns the thread count in effect."""
    try:
        cores = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        cores = os.cpu_count() or 1
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=cores, user_api='blas')
        n = [p.get('num_threads', 1) for p in threadpoolctl.threadpool_info() if p.get('user_api') == 'blas']
        return max(n) if n else cores
    except Exception:
        return cores


def load():
    """The CPU arm: the reference itself if it is installed under baseline/_ref, else the oracle port."""
    global _arm
    if _arm is not None:
        return _arm
    if os.path.isdir(os.path.join(REF_DIR, 'decomp')) and os.path.isdir(os.path.join(REF_DIR, 'chainer')):
        sys.path.insert(0, REF_DIR)
        try:
            import decomp as ref
            import decomp.dictionary_learning
            import decomp.lasso
            import decomp.nmf

            def lasso(y, A, alpha, **kw):
                return ref.lasso.solve(y, A, alpha, **kw)

            def nmf(y, D, **kw):
                return ref.nmf.solve(y, D, **kw)

            def dl(y, D, alpha, **kw):
                return ref.dictionary_learning.solve(y, D, alpha, **kw)

            _arm = CpuArm('reference', lasso, nmf, dl,
                          'unmodified reference (decomp %s) from baseline/_ref, numpy path' %
                          getattr(ref, '__version__', '?'))
            return _arm
        except Exception:
            sys.path.remove(REF_DIR)
    from oracle import decomp_oracle as orc
    _arm = CpuArm('port', orc.lasso, orc.nmf_mu, orc.dictionary_learning, 'oracle/decomp_oracle.py (numpy port)')
    return _arm
