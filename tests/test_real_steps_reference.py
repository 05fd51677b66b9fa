"""The harness of tests/test_gpu_real_steps.py with NOTHING patched: the reference's own step classes and functions on
the CPU reproduce the chain fixture.  This pins the harness (stub ``ORIGIN`` object, plain ``Cube`` / ``Image``
containers, the order and arguments of the step calls) independently of the B200 path it is then used to test."""

import pytest

from oracle import ref_loader
from real_steps_harness import check_against_fixture, run_pipeline

pytestmark = pytest.mark.skipif(ref_loader.find_reference_file('muse_origin/steps.py') is None,
                                reason='reference steps.py not present (oracle/_ref)')


def test_reference_steps_reproduce_the_chain_fixture(monkeypatch):
    g, out = run_pipeline(monkeypatch, 'reference')
    check_against_fixture(g, out)
