#!/usr/bin/env python
"""Install the UNMODIFIED reference (numpy path) into git-ignored ``baseline/_ref/`` so that it travels to the GPU
box with ``gpurun`` and ``bench.py --impl reference`` / the ``cpu_baseline`` legs can time the reference itself
(``kind: "reference"``) instead of the oracle port.  Nothing under ``baseline/_ref`` is tracked by git and the
product package never imports it.

    python tools/install_reference.py [--force]

Steps: (1) ``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref`` from a /tmp copy of
``/root/reference`` (the source tree is read-only and setuptools writes build files next to setup.py);
(2) the reference's setup.py lists only the top-level ``decomp`` package, so its sub-packages (``utils``,
``nmf_methods``, ``math_utils``) are not installed by pip: when the import check fails they are completed from the
source tree; (3) the reference hard-imports ``chainer`` (decomp/utils/cp_compat.py:1, decomp/template_matching.py:2)
although the hot path never uses it: an EMPTY 5-file stub package is written next to it.  Returns the outcome.
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get('DECOMP_REFERENCE', '/root/reference')
TARGET = os.path.join(ROOT, 'baseline', '_ref')
STUB = {
    'chainer/__init__.py': 'from . import cuda\n',
    'chainer/cuda.py': '',
    'chainer/utils/__init__.py': 'from . import conv, conv_nd\n',
    'chainer/utils/conv.py': '',
    'chainer/utils/conv_nd.py': '',
}


def _import_check():
    code = ('import sys; sys.path.insert(0, %r); import decomp; from decomp.nmf_methods import batch_mu; '
            'import decomp.lasso, decomp.nmf, decomp.dictionary_learning; print(decomp.__file__)' % TARGET)
    r = subprocess.run([sys.executable, '-c', code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return r.returncode == 0, r.stdout.strip()


def install(force=False):
    """Returns a one-line description of what happened ('' if the reference source is not available)."""
    if not os.path.isdir(os.path.join(REFERENCE, 'decomp')):
        return ''
    if os.path.isdir(os.path.join(TARGET, 'decomp')) and not force:
        ok, _ = _import_check()
        if ok:
            return 'baseline/_ref already holds an importable reference'
    if os.path.isdir(TARGET):
        shutil.rmtree(TARGET)
    os.makedirs(TARGET)
    how = []
    tmp = tempfile.mkdtemp(prefix='decomp_ref_src_')
    try:
        src = os.path.join(tmp, 'reference')
        shutil.copytree(REFERENCE, src, ignore=shutil.ignore_patterns('.git', '__pycache__'))
        cmd = [sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--no-deps',
               '--find-links', '/opt/wheelhouse', '--target', TARGET, src]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        how.append('pip install --target rc=%d' % r.returncode)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    for name, body in STUB.items():
        path = os.path.join(TARGET, name)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, 'w') as fh:
            fh.write(body)
    ok, msg = _import_check()
    if not ok:
        # sub-packages the reference's setup.py does not list: complete the installed package from the source tree
        dst = os.path.join(TARGET, 'decomp')
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(REFERENCE, 'decomp'), dst, ignore=shutil.ignore_patterns('__pycache__'))
        how.append('sub-packages completed from the source tree')
        ok, msg = _import_check()
    how.append('import ok: ' + msg if ok else 'IMPORT FAILED: ' + msg[-300:])
    return '; '.join(how)


if __name__ == '__main__':
    print(install(force='--force' in sys.argv))
