import os
import sys
import json

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def reference(module=None):
    """
    The UNMODIFIED reference package, imported from baseline/_ref (installed by baseline/install_reference.py; it travels to the
    GPU box) or, in the build container, from /root/reference.  Skips the calling test when neither exists.
    """
    import importlib
    base = os.path.join(ROOT, 'baseline')
    sys.path.insert(0, base)
    try:
        import install_reference
    finally:
        sys.path.remove(base)
    if install_reference.activate() is None:
        pytest.skip('reference not installed (run `python baseline/install_reference.py` where /root/reference exists)')
    sys.dont_write_bytecode = True
    return importlib.import_module('cosmoprimo' + ('.' + module if module else ''))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (run on the B200 box with -m gpu)')
    # the shared library is built in-tree; build it here if it is missing or stale (nvcc cross-compiles without a GPU)
    from cosmoprimo_b200 import _lib
    if _lib.needs_build():
        _lib.build()


class Golden(object):
    """Golden vectors written by tools/make_golden.py from the unmodified reference."""

    def __init__(self, name):
        self.data = np.load(os.path.join(GOLDEN_DIR, name), allow_pickle=False)
        self.cases = json.loads(str(self.data['manifest']))

    def inp(self, name):
        return self.data['in_' + name]

    def get(self, idx, name):
        return self.data['c{}_{}'.format(idx, name)]

    def has(self, idx, name):
        return 'c{}_{}'.format(idx, name) in self.data.files

    def fun(self, idx):
        case = self.cases[idx]
        return self.get(idx, 'fun') if case['inv'] else self.inp(case['fun'])

    @staticmethod
    def ctor_kwargs(case):
        return {k: (np.asarray(v) if k == 'q' and isinstance(v, list) else v) for k, v in case['ckw'].items()}

    @staticmethod
    def call_kwargs(case):
        kw = dict(case['callkw'])
        if isinstance(kw.get('extrap', None), list):
            kw['extrap'] = tuple(kw['extrap'])
        return kw


_golden_cache = {}


def load_golden(name='fftlog_golden.npz'):
    if name not in _golden_cache:
        _golden_cache[name] = Golden(name)
    return _golden_cache[name]


@pytest.fixture(scope='session')
def fftlog_golden():
    return load_golden('fftlog_golden.npz')


def scale_aware_error(G, G_ref, post):
    """max |dG| |w| / max |G_ref w| per row with w = 1/post (error in the biased space G y^q)."""
    w = 1. / np.abs(post)
    num = np.max(np.abs(np.asarray(G) - G_ref) * w, axis=-1)
    den = np.max(np.abs(G_ref) * w, axis=-1)
    return np.max(num / den)
