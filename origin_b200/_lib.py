"""ctypes binding of ``libogn.so`` (``include/ogn.h``).

There is no CPU fallback: if the shared library is missing or no B200 is
visible, every entry point raises.
"""

import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libogn.so')

OGN_F32, OGN_F64 = 0, 1
OGN_ERR_OVERFLOW = -5

c_void_p, c_int, c_int64, c_double = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double

# name -> (restype, argtypes); every symbol declared in include/ogn.h
SIGNATURES = {
    'ogn_version': (c_int, []),
    'ogn_create': (c_int, [c_int, c_void_p, ctypes.POINTER(c_void_p)]),
    'ogn_destroy': (None, [c_void_p]),
    'ogn_last_error': (ctypes.c_char_p, [c_void_p]),
    'ogn_synchronize': (c_int, [c_void_p]),
    'ogn_launch_count': (c_int64, [c_void_p]),
    'ogn_trim': (c_int, [c_void_p]),
    'ogn_fsf_folded': (c_int, [c_void_p, c_void_p]),
    'ogn_peer_alloc': (c_int, [c_void_p, ctypes.c_size_t, ctypes.POINTER(c_void_p), c_void_p]),
    'ogn_peer_free': (c_int, [c_void_p, c_void_p]),
    'ogn_peer_open': (c_int, [c_void_p, c_void_p, ctypes.POINTER(c_void_p)]),
    'ogn_peer_close': (c_int, [c_void_p, c_void_p]),
    'ogn_scatter_tile': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    'ogn_set_local_gather': (c_int, [c_void_p, c_void_p]),
    'ogn_peer_set_delay': (c_int, [c_void_p, c_int]),
    'ogn_peer_join': (c_int, [c_void_p]),
    'ogn_peer_sync': (c_int, [c_void_p]),
    'ogn_timing_enable': (c_int, [c_void_p, c_int]),
    'ogn_timing_report': (c_int, [c_void_p, ctypes.c_char_p, ctypes.c_size_t]),
    'ogn_variants': (c_int, [c_void_p, ctypes.c_char_p, ctypes.c_size_t]),
    'ogn_host_alloc': (c_int, [ctypes.c_size_t, ctypes.POINTER(c_void_p)]),
    'ogn_host_free': (c_int, [c_void_p]),
    'ogn_tglr': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                         c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'ogn_fsf_stage': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                              c_void_p, c_void_p]),
    'ogn_step05': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                           c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                           c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                           c_void_p]),
    'ogn_step05_bits': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                                c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    'ogn_step05_tile': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int,
                                c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                c_void_p]),
    'ogn_local_extrema': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    'ogn_purity_stats': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p,
                                 c_int, c_int, c_void_p, c_void_p]),
    'ogn_purity_counts': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p,
                                  c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    'ogn_purity_counts_dev': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                      c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    'ogn_threshold_extract': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_double, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_int64, c_void_p]),
    'ogn_dct_residual': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                 c_void_p, c_int]),
    'ogn_preprocess_begin': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                     c_int, c_void_p, c_void_p, c_void_p]),
    'ogn_preprocess_finish': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p]),
    'ogn_greedy_pca': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p, c_int64, c_void_p, c_double, c_double,
                               c_int, c_void_p, c_int, c_void_p, c_void_p]),
    'ogn_line_estimates': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                                   c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'ogn_line_estimates_fields': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int,
                                          c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    'ogn_preprocess': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
}

_lib = None
_lock = threading.Lock()


class OgnError(RuntimeError):
    def __init__(self, code, message):
        super().__init__('libogn error %d: %s' % (code, message))
        self.code = code


def load_library(path=None):
    """Load libogn.so and declare every prototype.  Raises if it is absent."""
    global _lib
    with _lock:
        if _lib is not None and path is None:
            return _lib
        p = path or LIB_PATH
        if not os.path.exists(p):
            raise OSError('%s not found: build it with `python -m origin_b200.build` '
                          '(there is no CPU fallback for the ported path)' % p)
        lib = ctypes.CDLL(p)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if a declared symbol is missing
            fn.restype = restype
            fn.argtypes = argtypes
        if path is None:
            _lib = lib
        return lib


def _is_torch(x):
    return type(x).__module__.startswith('torch')


def ptr(x):
    """Raw address of a numpy array / torch tensor (host or device), or None."""
    if x is None:
        return None
    if _is_torch(x):
        if not x.is_contiguous():
            raise ValueError('tensor must be contiguous')
        return x.data_ptr()
    if not x.flags['C_CONTIGUOUS']:
        raise ValueError('array must be C-contiguous')
    return x.ctypes.data


class Context:
    """One libogn context: (device, stream).  Not thread-safe."""

    def __init__(self, device=0, stream=None):
        self.lib = load_library()
        handle = c_void_p()
        rc = self.lib.ogn_create(int(device), c_void_p(stream or 0), ctypes.byref(handle))
        if rc != 0:
            raise OgnError(rc, (self.lib.ogn_last_error(None) or b'').decode())
        self.handle = handle
        self.device = int(device)
        self.stream = stream or 0

    def check(self, rc, allow_overflow=False):
        if rc == 0 or (allow_overflow and rc == OGN_ERR_OVERFLOW):
            return rc
        raise OgnError(rc, (self.lib.ogn_last_error(self.handle) or b'').decode())

    def synchronize(self):
        self.check(self.lib.ogn_synchronize(self.handle))

    @property
    def launch_count(self):
        return int(self.lib.ogn_launch_count(self.handle))

    def timing(self, on=True):
        """Switch per-stage CUDA-event timing on or off."""
        self.check(self.lib.ogn_timing_enable(self.handle, int(bool(on))))

    def timing_report(self):
        """``[(stage, ms), ...]`` for the stages timed since the last report."""
        buf = ctypes.create_string_buffer(1 << 20)
        self.check(self.lib.ogn_timing_report(self.handle, buf, len(buf)))
        out = []
        for item in buf.value.decode().split(';'):
            if item:
                name, ms = item.rsplit(':', 1)
                out.append((name, float(ms)))
        return out

    def variants(self):
        """``{stage: code path}`` of the last launch of each stage on this context (``ogn_variants``)."""
        buf = ctypes.create_string_buffer(4096)
        self.check(self.lib.ogn_variants(self.handle, buf, len(buf)))
        return dict(item.split('=', 1) for item in buf.value.decode().split(';') if item)

    def trim(self):
        self.check(self.lib.ogn_trim(self.handle))

    @property
    def fsf_folded(self):
        """True when the last TGLR call used the row-folded spatial kernel (mirror-symmetric FSF)."""
        v = c_int(0)
        self.check(self.lib.ogn_fsf_folded(self.handle, ctypes.byref(v)))
        return bool(v.value)

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.ogn_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=None):
    """Process-wide context on ``device`` (default: the current torch device, or
    0) and the legacy default stream."""
    if device is None:
        device = 0
        try:
            import torch
            if torch.cuda.is_available():
                device = torch.cuda.current_device()
        except ImportError:
            pass
    ctx = _default_ctx.get(device)
    if ctx is None or ctx.handle is None:
        ctx = _default_ctx[device] = Context(device)
    return ctx


def pinned_empty(shape, dtype):
    """numpy array backed by page-locked host memory (fast H2D/D2H staging)."""
    lib = load_library()
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    p = c_void_p()
    rc = lib.ogn_host_alloc(max(nbytes, 16), ctypes.byref(p))
    if rc != 0:
        raise OgnError(rc, (lib.ogn_last_error(None) or b'').decode())
    buf = (ctypes.c_char * max(nbytes, 16)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PinnedOwner(arr, p.value, lib)
    return arr


class _PinnedOwner:
    """Frees the pinned block when the numpy view dies."""
    _live = {}

    def __init__(self, arr, address, lib):
        import weakref
        self.address, self.lib = address, lib
        _PinnedOwner._live[address] = self
        weakref.finalize(arr.base if arr.base is not None else arr, self._free)

    def _free(self):
        if _PinnedOwner._live.pop(self.address, None) is not None:
            self.lib.ogn_host_free(c_void_p(self.address))
