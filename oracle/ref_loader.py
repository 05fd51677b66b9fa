"""Load the reference's ``lib_origin.py`` UNMODIFIED, by path.  TEST / BENCH INFRASTRUCTURE ONLY.

``import muse_origin`` fails in this image (no generated ``version.py``; astropy, mpdaf, photutils and
matplotlib are not installed), but the hot-path functions only touch numpy / scipy / joblib.
Registering stub modules for the missing imports lets the file execute as is (SURVEY.md §8c /
appendix A.1).  Search order: ``oracle/_ref`` (the travelling copy made by ``oracle/build_ref.py``),
then ``/root/reference`` (build container).  Used by ``tests/golden/make_golden.py`` to produce the
committed fixtures, by the CPU tests that pin signatures against the real step layer, and by
``bench.py`` for the CPU baseline; never by the product.
"""

import importlib.util
import os
import sys
import types
from unittest.mock import MagicMock

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get('ORIGIN_REFERENCE_ROOT', '/root/reference')


def reference_roots():
    return [os.path.join(HERE, '_ref'), REFERENCE_ROOT]


def find_reference_file(rel):
    for root in reference_roots():
        p = os.path.join(root, rel)
        if os.path.exists(p):
            return p
    return None


def available():
    return find_reference_file('muse_origin/lib_origin.py') is not None


class TableShim:
    """The few lines of ``astropy.table.Table`` that ``Compute_threshold_purity``
    (lib_origin.py:1454-1470) touches."""

    class _Col:
        def __init__(self, data):
            import numpy as np
            self.data = np.asarray(data)
            self.format = None

        def __array__(self, dtype=None, copy=None):
            return self.data if dtype is None else self.data.astype(dtype)

        def __len__(self):
            return len(self.data)

        def __getitem__(self, i):
            return self.data[i]

    def __init__(self, cols, names):
        self.names = list(names)
        self.cols = {n: self._Col(c) for n, c in zip(names, cols)}

    def __getitem__(self, name):
        return self.cols[name]

    def sort(self, key):
        import numpy as np
        order = np.argsort(self.cols[key].data, kind='stable')
        for c in self.cols.values():
            c.data = c.data[order]

    def __str__(self):
        return 'TableShim(%s)' % ', '.join(self.names)


class _ProgressBar:
    """``mpdaf.tools.progressbar`` (a tqdm wrapper) reduced to what the hot path touches: iteration over an iterable
    (``Correlation_GLR_test``, ``Compute_threshold_purity``) and the context-manager form with ``update`` / ``n``
    (``Compute_GreedyPCA``, lib_origin.py:884-950)."""

    def __init__(self, iterable=None, total=None, **kw):
        self.iterable, self.total, self.n = iterable, total, 0

    def __iter__(self):
        return iter(self.iterable)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def update(self, k=1):
        self.n += k


STUBS = ['matplotlib', 'matplotlib.pyplot', 'astropy', 'astropy.modeling', 'astropy.modeling.fitting',
         'astropy.modeling.models', 'astropy.nddata', 'astropy.stats', 'astropy.table', 'astropy.utils',
         'astropy.utils.exceptions', 'astropy.io', 'astropy.io.fits', 'astropy.units', 'mpdaf', 'mpdaf.obj',
         'mpdaf.tools', 'mpdaf.log', 'mpdaf.MUSE', 'mpdaf.sdetect', 'photutils', 'yaml']


def _install_stubs():
    for name in STUBS:
        if name not in sys.modules:
            try:
                if name == 'yaml':
                    import yaml  # noqa: F401  (present in this image; stub only if missing)
                    continue
            except ImportError:
                pass
            sys.modules[name] = MagicMock()
    sys.modules['mpdaf.tools'].progressbar = _ProgressBar

    class AstropyUserWarning(Warning):
        pass

    sys.modules['astropy.utils.exceptions'].AstropyUserWarning = AstropyUserWarning
    if 'muse_origin' not in sys.modules or not isinstance(sys.modules['muse_origin'], types.ModuleType) \
            or not hasattr(sys.modules['muse_origin'], '__ogn_stub__'):
        pkg = types.ModuleType('muse_origin')
        pkg.__path__ = [os.path.dirname(find_reference_file('muse_origin/lib_origin.py'))]
        pkg.__ogn_stub__ = True
        pkg.__version__ = 'reference'
        sys.modules['muse_origin'] = pkg
        sys.modules['muse_origin.source_masks'] = MagicMock()
        sys.modules['muse_origin.version'] = types.ModuleType('muse_origin.version')
        sys.modules['muse_origin.version'].version = 'reference'


def _load(modname, rel):
    if modname in sys.modules and getattr(sys.modules[modname], '__ogn_loaded__', False):
        return sys.modules[modname]
    path = find_reference_file(rel)
    if path is None:
        raise RuntimeError('reference file %s not present (looked in %s)' % (rel, ', '.join(reference_roots())))
    _install_stubs()
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    mod.__ogn_loaded__ = True
    return mod


def load_lib_origin():
    """The reference's ``muse_origin.lib_origin`` module, executed unmodified."""
    lib = _load('muse_origin.lib_origin', 'muse_origin/lib_origin.py')
    lib.Table = TableShim
    return lib


def load_steps():
    """The reference's ``muse_origin.steps`` module, executed unmodified with mpdaf / astropy stubbed: its ``Step``
    classes, ``DataObj`` descriptors and ``dump`` run as they are once ``Cube`` / ``Image`` are replaced by plain
    containers (tests/test_gpu_real_steps.py)."""
    load_lib_origin()
    return _load('muse_origin.steps', 'muse_origin/steps.py')
