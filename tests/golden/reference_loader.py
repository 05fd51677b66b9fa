"""Kept for ``make_golden.py``: the loader now lives in ``oracle/ref_loader.py`` (it also serves the
CPU baseline of ``bench.py`` through the travelling copy ``oracle/_ref``)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.ref_loader import REFERENCE_ROOT, TableShim, load_lib_origin, load_steps  # noqa: E402,F401
