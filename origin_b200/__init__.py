"""origin_b200 — B200-native (sm_100a) implementation of ORIGIN's detection hot path.

Host layer mirroring the reference's interfaces for the path
step01 DCT -> step05 TGLR -> local extrema -> step06 purity -> step07 thresholding,
on top of the C-ABI library ``libogn.so`` (``include/ogn.h``).
"""

from . import dictionaries, synthetic  # noqa: F401

__version__ = '0.1.0'


def __getattr__(name):
    # lib_origin / steps / distributed need libogn only when used; import lazily
    if name in ('lib_origin', 'steps', 'tiles', 'distributed'):
        import importlib
        return importlib.import_module('.' + name, __name__)
    raise AttributeError(name)
