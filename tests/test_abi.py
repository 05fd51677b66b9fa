"""The C-ABI library builds, loads and exports every symbol ``include/ogn.h``
declares (no compute calls: this runs without a GPU)."""

import ctypes
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope='module')
def libpath():
    from origin_b200 import build
    return build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'ogn.h')).read()
    return sorted(set(re.findall(r'OGN_API[^;(]*?\b(ogn_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_expected_entry_points():
    names = declared_symbols()
    for must in ('ogn_create', 'ogn_destroy', 'ogn_last_error', 'ogn_tglr', 'ogn_local_extrema',
                 'ogn_purity_counts', 'ogn_threshold_extract', 'ogn_dct_residual',
                 'ogn_preprocess_begin', 'ogn_preprocess_finish'):
        assert must in names


def test_library_exports_every_declared_symbol(libpath):
    lib = ctypes.CDLL(libpath)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_binding_table_matches_header(libpath):
    from origin_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load_library()
    assert lib.ogn_version() == 100


def test_no_cpu_fallback_without_device(libpath):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from origin_b200 import _lib
    with pytest.raises(_lib.OgnError) as err:
        _lib.Context(0)
    assert 'no CPU fallback' in str(err.value)
    from origin_b200 import lib_origin
    import numpy as np
    with pytest.raises(_lib.OgnError):
        lib_origin.compute_local_max(np.zeros((3, 3, 3), np.float32), np.zeros((3, 3, 3), np.float32),
                                     np.zeros((3, 3, 3), bool))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'origin_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', text, re.M), os.path.join(dirpath, f)
                assert 'origin_oracle' not in text and 'oracle/' not in text, os.path.join(dirpath, f)
