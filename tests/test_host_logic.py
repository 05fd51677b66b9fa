"""Host-side logic that needs no GPU: the step06 tail against the reference's own function on the edge
cases it handles specially, and the signatures of the drop-in layer against the reference's step layer."""

import inspect

import numpy as np
import pytest

from oracle import ref_loader
from test_tiles_dist import NumpyCounter, _fake_extrema

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason='reference module not present (oracle/_ref)')


def _lists(lmax, lmin):
    from origin_b200 import lib_origin
    mi, ni = np.flatnonzero(lmax), np.flatnonzero(lmin)
    return lib_origin.LocalExtrema(lmax.shape, mi, lmax.reshape(-1)[mi], ni, lmin.reshape(-1)[ni])


@needs_ref
@pytest.mark.parametrize('case', ['default', 'segmap', 'negative', 'negative_segmap', 'decreasing_default', 'unsorted'])
def test_threshold_purity_tail_matches_the_reference(case):
    """``Compute_threshold_purity`` (lib_origin.py:1391-1479) on dense cubes vs our list-based tail, including
    thresholds below zero (zeros of the dense cubes count, background-only with a segmap, :1429) and a
    decreasing default list (threshmax < threshmin: the '> threshmin' prefilter of :1443 decides)."""
    import warnings
    from origin_b200 import lib_origin
    lib = ref_loader.load_lib_origin()
    lmax, lmin, seg = _fake_extrema(11, (10, 14, 12))
    lmax, lmin = lmax.astype(np.float64), lmin.astype(np.float64)
    segmap, thr = None, None
    if case in ('segmap', 'negative_segmap'):
        segmap = seg
    if case.startswith('negative'):
        thr = np.linspace(-2.0, 6.0, 17)
    if case == 'unsorted':
        thr = np.array([5.0, 1.0, 3.0, 2.0, 8.0])
    if case == 'decreasing_default':
        lmin = lmin * 0.05                       # max(lmin) far below 1.1 * median(max_z lmax)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        rthr, rtab = lib.Compute_threshold_purity(0.6, lmax.copy(), lmin.copy(), segmap, thr)
        thr_, tab = lib_origin.Compute_threshold_purity(0.6, _lists(lmax.astype(np.float32), lmin.astype(np.float32)),
                                                        None, segmap, thr, _backend=NumpyCounter())
    np.testing.assert_allclose(tab['Tval_r'], np.asarray(rtab['Tval_r']), rtol=1e-6)
    np.testing.assert_array_equal(tab['Det_M'], np.asarray(rtab['Det_M']))
    np.testing.assert_array_equal(tab['Det_m'], np.asarray(rtab['Det_m']))
    assert (np.isinf(thr_) and np.isinf(rthr)) or thr_ == pytest.approx(rthr, rel=1e-6)


def test_check_counts_flags_the_overflow_mark():
    from origin_b200 import lib_origin
    lib_origin.check_counts(np.array([3, 0]), np.array([1, 0]))
    with pytest.raises(OverflowError):
        lib_origin.check_counts(np.array([3, 0]), np.array([-(1 << 56) + 5, 2]))


@pytest.mark.skipif(ref_loader.find_reference_file('muse_origin/steps.py') is None,
                    reason='reference steps.py only exists in the build container')
def test_drop_in_signatures_equal_the_reference():
    """The fused ``run`` methods patched into the reference's step classes and the function mirror keep the
    reference's signatures (steps.py:420-429, 756, 851-859; lib_origin.py:150, 1070, 1220, 1391)."""
    from origin_b200 import lib_origin, steps as osteps
    rsteps = ref_loader.load_steps()
    lib = ref_loader.load_lib_origin()

    def params(fn):
        return [(p.name, p.default) for p in inspect.signature(fn).parameters.values()]

    assert params(osteps._run_preprocessing) == params(rsteps.Preprocessing.run)
    assert params(osteps._run_compute_tglr) == params(rsteps.ComputeTGLR.run)
    assert params(osteps._run_purity) == params(rsteps.ComputePurityThreshold.run)
    for name in ('dct_residual', 'Correlation_GLR_test', 'compute_local_max', 'Compute_threshold_purity', 'O2test',
                 'DCTMAT'):
        ref = params(getattr(lib, name))
        ours = params(getattr(lib_origin, name))
        assert ours[:len(ref)] == ref, name           # ours may append keyword-only extras (ctx, out_dtype)
    # what the step layer imports by name (steps.py:19-41) is what patch_steps rebinds
    for name in ('dct_residual', 'compute_local_max', 'Correlation_GLR_test', 'Compute_threshold_purity', 'O2test'):
        assert hasattr(rsteps, name)


@pytest.mark.skipif(ref_loader.find_reference_file('muse_origin/steps.py') is None,
                    reason='reference steps.py only exists in the build container')
def test_step04_step08_drop_ins_keep_the_reference_signatures():
    """``ComputeGreedyPCA.run`` (steps.py:681), ``Compute_GreedyPCA_area`` (lib_origin.py:769) and
    ``estimation_line`` (:1805) as ``patch_steps`` rebinds them."""
    from origin_b200 import lib_origin, steps as osteps
    rsteps = ref_loader.load_steps()
    lib = ref_loader.load_lib_origin()

    def params(fn):
        return [(p.name, p.default) for p in inspect.signature(fn).parameters.values()]

    assert params(osteps._run_greedy_pca) == params(rsteps.ComputeGreedyPCA.run)
    assert params(osteps._run_pca_threshold) == params(rsteps.ComputePCAThreshold.run)
    assert params(osteps._pca_threshold) == params(lib.Compute_PCA_threshold)
    assert params(osteps._estimation_line_table) == params(lib.estimation_line)
    ref = params(lib.Compute_GreedyPCA_area)
    assert params(lib_origin.Compute_GreedyPCA_area)[:len(ref)] == ref
    for name in ('Compute_PCA_threshold', 'Compute_GreedyPCA_area', 'estimation_line'):
        assert hasattr(rsteps, name)                      # imported by name at steps.py:19-41


def test_cat2_table_inserts_the_columns_where_the_reference_does():
    """An astropy-like ``Cat1`` comes back as a copy with ``ra / dec / lbda`` updated and ``x, y, z, residual, flux,
    num_line`` inserted at indexes 4, 5, 6, 8, 8, 8 of the ORIGINAL column list (lib_origin.py:1925-1936)."""
    from origin_b200 import steps as osteps

    class FakeColumn:
        def __init__(self, name, data):
            self.name, self.data = name, np.asarray(data)

    class Table:
        Column = FakeColumn                              # astropy's Table carries its Column class the same way

        def __init__(self, cols):
            self.cols = dict(cols)

        def copy(self):
            return Table(self.cols)

        def __setitem__(self, key, val):
            self.cols[key] = np.asarray(val)

        def add_columns(self, cols, indexes):
            names = list(self.cols)
            merged = []
            for i, n in enumerate(names + [None]):       # astropy: each new column goes BEFORE original index i
                merged += [c.name for c, at in zip(cols, indexes) if at == i]
                if n is not None:
                    merged.append(n)
            new = {c.name: c.data for c in cols}
            self.cols = {n: new.get(n, self.cols.get(n)) for n in merged}

    cat1 = Table({k: np.arange(2) for k in ('ra', 'dec', 'lbda', 'x0', 'y0', 'z0', 'comp', 'STD', 'T_GLR', 'profile')})
    cat2 = dict(ra=np.array([1.5, 2.5]), dec=np.array([3.0, 4.0]), lbda=np.array([5000.0, 6000.0]), x=np.array([7, 8]),
                y=np.array([9, 10]), z=np.array([11, 12]), residual=np.array([0.1, 0.2]), flux=np.array([20.0, 30.0]),
                num_line=np.array([1, 2]))
    out = osteps._cat2_table(cat1, cat2)
    assert out is not cat1 and list(cat1.cols) == ['ra', 'dec', 'lbda', 'x0', 'y0', 'z0', 'comp', 'STD', 'T_GLR', 'profile']
    # the reference's Cat2 column order (docstring of estimation_line, :1853-1855, with comp / STD kept)
    assert list(out.cols) == ['ra', 'dec', 'lbda', 'x0', 'x', 'y0', 'y', 'z0', 'z', 'comp', 'STD', 'residual', 'flux',
                              'num_line', 'T_GLR', 'profile']
    np.testing.assert_array_equal(out.cols['ra'], cat2['ra'])
    np.testing.assert_array_equal(out.cols['flux'], cat2['flux'])
    assert osteps._cat2_table(dict(x0=[1]), cat2) is cat2


def test_patch_steps_rebinds_and_restores_names_that_were_absent():
    import types
    from origin_b200 import lib_origin, steps as osteps
    mod = types.ModuleType('stand_in')
    mod.dct_residual = sentinel = object()
    osteps.patch_steps(mod, fused=True)                  # no step classes in this stand-in: only the names
    try:
        assert mod.estimation_line is osteps._estimation_line_table
        assert mod.Compute_GreedyPCA_area is lib_origin.Compute_GreedyPCA_area
    finally:
        osteps.unpatch_steps()
    assert mod.dct_residual is sentinel and not hasattr(mod, 'estimation_line') and not hasattr(mod, 'O2test')


@pytest.mark.skipif(ref_loader.find_reference_file('muse_origin/steps.py') is None,
                    reason='reference steps.py only exists in the build container')
def test_lazy_products_live_in_the_reference_step_machinery(tmp_path):
    """``patch_steps`` on the REAL ``muse_origin.steps`` module (stub-loaded: mpdaf / astropy are mocks) and a
    ``LazyProduct`` going through the reference's own ``DataObj`` descriptor (steps.py:121-164), ``store_cube``
    (:284-294) and ``Step.dump`` (:301-337): the placeholder is handed out untouched, ``dump`` makes it fetch, writes
    the float64 cube with ``convert_float32=False`` and leaves the file path behind like for any other product."""
    from origin_b200 import lib_origin, steps as osteps
    rsteps = ref_loader.load_steps()
    originals = {c: getattr(rsteps, c).run for c in ('Preprocessing', 'ComputeGreedyPCA', 'ComputeTGLR', 'ComputePurityThreshold')}
    ref_names = {n: getattr(rsteps, n) for n in ('Correlation_GLR_test', 'estimation_line', 'Compute_GreedyPCA_area')}
    written = []

    class FakeCube:
        def __init__(self, data=None, **kw):
            self.data, self.kw = data, kw

        def write(self, path, **kw):
            written.append((path, kw, self.data))

    class Orig:
        wave = wcs = None
        steps = {}

    real_cube = rsteps.Cube
    osteps.patch_steps(rsteps, fused=True)
    try:
        assert rsteps.ComputeTGLR.run is osteps._run_compute_tglr and rsteps.ComputeGreedyPCA.run is osteps._run_greedy_pca
        assert rsteps.Correlation_GLR_test is lib_origin.Correlation_GLR_test
        assert rsteps.estimation_line is osteps._estimation_line_table
        rsteps.Cube = FakeCube
        step = rsteps.ComputeTGLR(Orig(), 5, {})
        fetched = []

        def fetch():
            fetched.append(1)
            return np.arange(24, dtype=np.float64).reshape(2, 3, 4)

        step.cube_correl_min = osteps.LazyProduct(step, 'cube_correl_min', fetch, 'cube', (2, 3, 4))     # DataObj.__set__
        assert isinstance(step.cube_correl_min, osteps.LazyProduct) and step.cube_correl_min.shape == (2, 3, 4)
        assert step.cube_correl_min.on_device() is None and not fetched                                  # DataObj.__get__
        step.status = rsteps.Status.RUN
        step.dump(str(tmp_path))
        assert fetched == [1] and len(written) == 1
        path, kw, data = written[0]
        assert path.endswith('cube_correl_min.fits') and kw == dict(convert_float32=False) and data.dtype == np.float64
        assert step.__dict__['cube_correl_min'] == path and step.status is rsteps.Status.DUMPED
        # reading a product materialises it once and replaces the placeholder by the step's own Cube
        step.cube_profile = osteps.LazyProduct(step, 'cube_profile', lambda: np.zeros((2, 3, 4), np.uint8), 'cube', (2, 3, 4))
        assert step.cube_profile.data.dtype == np.uint8 and isinstance(step.__dict__['cube_profile'], FakeCube)
    finally:
        rsteps.Cube = real_cube
        osteps.unpatch_steps()
    for c, run in originals.items():
        assert getattr(rsteps, c).run is run
    for n, fn in ref_names.items():
        assert getattr(rsteps, n) is fn


@pytest.mark.skipif(ref_loader.find_reference_file('muse_origin/steps.py') is None,
                    reason='reference steps.py not present (oracle/_ref)')
def test_fused_pca_threshold_reads_the_o2_map_of_step01():
    """Fused step03 on the reference's real ``ComputePCAThreshold`` class (steps.py:572-631): the O2 test of each area
    comes from the map the fused step01 left (only when it was made from the very cube the step holds), otherwise it
    is recomputed from the cube like ``Compute_PCA_threshold`` (lib_origin.py:821-842) does; both give what the
    rebound function gives area by area."""
    from origin_b200 import steps as osteps
    rsteps = ref_loader.load_steps()
    rng = np.random.default_rng(5)
    cube = rng.standard_normal((60, 24, 30)).astype(np.float32)
    cube[:, 3:6, 4:9] *= 1.6
    areamap = np.ones((24, 30), dtype=int)
    areamap[:, 15:] = 2
    o2 = np.mean(cube.astype(np.float64) ** 2, axis=0)

    class Holder:
        def __init__(self, data):
            self._data = data

    class Orig:
        pass

    osteps.patch_steps(rsteps, fused=True)
    try:
        assert rsteps.ComputePCAThreshold.run is osteps._run_pca_threshold
        assert rsteps.Compute_PCA_threshold is osteps._pca_threshold
        expect = [osteps._pca_threshold(cube[:, areamap == a], 0.01) for a in (1, 2)]
        for cached, marker in (((cube, o2), 1.0), ((cube, o2 * 1.5), 1.5), ((cube.copy(), o2 * 1.5), 1.0), (None, 1.0)):
            orig = Orig()
            orig.cube_std, orig.areamap, orig.nbAreas = Holder(cube), Holder(areamap), 2
            pre = rsteps.Preprocessing(orig, 1, {})
            if cached is not None:
                pre._ogn_o2map = cached
            orig.steps = {'preprocessing': pre}
            step = rsteps.ComputePCAThreshold(orig, 3, {})
            step.run(orig)
            assert len(step.thresO2) == len(orig.testO2) == len(orig.histO2) == len(orig.binO2) == 2
            for a in range(2):
                # marker 1.5: the (deliberately scaled) cached map was used; 1.0: the cube was reduced again
                np.testing.assert_allclose(orig.testO2[a], expect[a][0] * marker, rtol=1e-12)
            if marker == 1.0:
                np.testing.assert_allclose(step.thresO2, [e[3] for e in expect], rtol=1e-10)
                np.testing.assert_allclose(step.meaO2, [e[4] for e in expect], rtol=1e-10)
    finally:
        osteps.unpatch_steps()
