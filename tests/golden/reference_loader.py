"""Load the reference's ``lib_origin.py`` UNMODIFIED, by path (build container only).

``import muse_origin`` fails here (no generated ``version.py``; astropy, mpdaf,
photutils and matplotlib are not installed), but the hot-path functions only
touch numpy / scipy / joblib.  Registering stub modules for the missing
imports lets ``/root/reference/muse_origin/lib_origin.py`` execute as is
(SURVEY.md §8c / appendix A.1).  Used by ``make_golden.py`` to produce the
fixtures in this directory; never imported on the GPU box (the reference does
not travel) and never by the product.
"""

import importlib.util
import os
import sys
import types
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get('ORIGIN_REFERENCE_ROOT', '/root/reference')


class TableShim:
    """The 15 lines of ``astropy.table.Table`` that
    ``Compute_threshold_purity`` (lib_origin.py:1454-1470) touches."""

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


def load_lib_origin():
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError('reference tree not present: %s' % REFERENCE_ROOT)
    if 'muse_origin.lib_origin' in sys.modules:
        return sys.modules['muse_origin.lib_origin']
    stubs = ['matplotlib', 'matplotlib.pyplot', 'astropy', 'astropy.modeling',
             'astropy.modeling.fitting', 'astropy.modeling.models', 'astropy.nddata',
             'astropy.stats', 'astropy.table', 'astropy.utils',
             'astropy.utils.exceptions', 'astropy.io', 'astropy.units', 'mpdaf',
             'mpdaf.obj', 'mpdaf.tools', 'photutils']
    for name in stubs:
        if name not in sys.modules:
            sys.modules[name] = MagicMock()
    sys.modules['mpdaf.tools'].progressbar = lambda it=None, **kw: it

    class AstropyUserWarning(Warning):
        pass

    sys.modules['astropy.utils.exceptions'].AstropyUserWarning = AstropyUserWarning
    pkg = types.ModuleType('muse_origin')
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, 'muse_origin')]
    sys.modules['muse_origin'] = pkg
    sys.modules['muse_origin.source_masks'] = MagicMock()
    spec = importlib.util.spec_from_file_location(
        'muse_origin.lib_origin', os.path.join(REFERENCE_ROOT, 'muse_origin', 'lib_origin.py'))
    lib = importlib.util.module_from_spec(spec)
    sys.modules['muse_origin.lib_origin'] = lib
    spec.loader.exec_module(lib)
    lib.Table = TableShim
    return lib
