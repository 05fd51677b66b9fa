"""Host-side eigen-solver of the Lanczos users (``ogn_lz::tridiag_top``, origin_b200/csrc/ogn_lanczos.cuh: Sturm
bisection for the largest eigenvalue of a symmetric tridiagonal matrix, inverse iteration with a pivoted tridiagonal
LU for its eigenvector) against ``scipy.linalg.eigh_tridiagonal``.  The step04 greedy PCA and the step08 line
estimation take their singular vectors from it on every restart cycle; it is plain host C++, so it is compiled into a
small harness with nvcc and runs here without a GPU."""

import os
import shutil
import subprocess

import numpy as np
import pytest
from scipy.linalg import eigh_tridiagonal

from conftest import ROOT

HARNESS = r'''
#include <cstdio>
#include <vector>
#include "ogn_lanczos.cuh"
int main() {
    int k;
    while (scanf("%d", &k) == 1) {
        std::vector<double> a(k), b(k > 1 ? k : 1, 0.0), y;
        for (int i = 0; i < k; ++i) scanf("%lf", &a[i]);
        for (int i = 0; i + 1 < k; ++i) scanf("%lf", &b[i]);
        double theta = 0.0;
        ogn_lz::tridiag_top(a, b, k, &theta, &y);
        printf("%.17g", theta);
        for (int i = 0; i < k; ++i) printf(" %.17g", y[i]);
        printf("\n");
    }
    return 0;
}
'''


@pytest.fixture(scope='module')
def solver(tmp_path_factory):
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        pytest.skip('nvcc not found')
    d = tmp_path_factory.mktemp('tridiag')
    src = d / 'harness.cu'
    src.write_text(HARNESS)
    exe = d / 'harness'
    subprocess.run([nvcc, '-O2', '-I', os.path.join(ROOT, 'origin_b200', 'csrc'), '-o', str(exe), str(src)], check=True,
                   capture_output=True)

    def solve(cases):
        text = ''.join('%d\n%s\n%s\n' % (len(a), ' '.join('%.17g' % v for v in a), ' '.join('%.17g' % v for v in b))
                       for a, b in cases)
        out = subprocess.run([str(exe)], input=text, capture_output=True, text=True, check=True).stdout.splitlines()
        res = [np.array(line.split(), dtype=np.float64) for line in out]
        return [(r[0], r[1:]) for r in res]
    return solve


def cases():
    rng = np.random.default_rng(12)
    out = []
    for k in (2, 3, 5, 16, 24, 40, 48):                        # the Krylov sizes in use and around them
        out.append((rng.normal(size=k), rng.normal(size=k - 1)))
        # what Lanczos on X X^T produces: positive diagonal, positive off-diagonals decaying fast, a dominant first entry
        a = np.abs(rng.normal(size=k)) * np.logspace(0, -6, k) * 1e4
        b = np.abs(rng.normal(size=k - 1)) * np.logspace(0, -7, k - 1) * 1e3
        out.append((a, b))
    out.append((np.array([3.0, 3.0, 3.0, 3.0]), np.array([1e-9, 1e-9, 1e-9])))       # nearly degenerate
    out.append((np.array([1.0, 5.0, 2.0]), np.array([0.0, 0.0])))                      # decoupled: the top pair sits in the middle
    out.append((np.array([1e-150, 2e-150, 3e-150]), np.array([1e-151, 1e-151])))       # tiny scale
    out.append((np.array([1e150, 2e150]), np.array([5e149])))                           # huge scale
    out.append((np.array([0.0, 0.0, 0.0, 0.0, 0.0]), np.array([1.0, 1.0, 1.0, 1.0])))  # zero diagonal (pivoting in the LU)
    out.append((np.array([2.0, 1.0]), np.array([1e3])))                                 # off-diagonal dominates: row swaps
    return out


def test_top_eigenpair_matches_scipy(solver):
    cs = cases() + [(np.array([7.5]), np.array([]))]
    got = solver(cs)
    assert len(got) == len(cs)
    for (a, b), (theta, y) in zip(cs, got):
        k = len(a)
        if k == 1:
            assert theta == a[0] and y[0] == 1.0
            continue
        w, v = eigh_tridiagonal(a, b)
        scale = max(np.abs(w).max(), 1e-300)
        assert abs(theta - w[-1]) <= 4e-15 * scale, (k, theta, w[-1])
        assert abs(np.linalg.norm(y) - 1.0) <= 1e-12
        # residual of the pair in the matrix norm (eigenvectors of clustered eigenvalues are not unique; the residual is)
        t = np.diag(a) + np.diag(b, 1) + np.diag(b, -1)
        assert np.linalg.norm(t @ y - theta * y) <= 1e-13 * scale * np.sqrt(k), (k, a[:3], b[:3])
        gap = w[-1] - w[-2]
        if gap > 1e-6 * scale:                                   # well separated: the vector itself (up to sign)
            assert min(np.abs(y - v[:, -1]).max(), np.abs(y + v[:, -1]).max()) <= 1e-9
