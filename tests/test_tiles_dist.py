"""Host-side logic of the multi-GPU path on CPU: tile planning, index mapping and the
world_size-2 gloo reductions of the purity counts (the kernels are replaced by the numpy
oracle through the ``_backend`` hook; the GPU twin is ``tools/check_sharded.py``)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import origin_oracle as orc
from origin_b200 import tiles


@pytest.mark.parametrize('n,ny,nx', [(1, 40, 50), (2, 320, 320), (4, 65, 80), (8, 320, 320), (8, 900, 900), (3, 31, 17)])
def test_plan_tiles_partitions_the_field(n, ny, nx):
    plan = tiles.plan_tiles(ny, nx, n, 13)
    assert len(plan) == n
    cover = np.zeros((ny, nx), dtype=int)
    for t in plan:
        cover[t.global_owned] += 1
        assert t.py0 == max(0, t.y0 - 13) and t.py1 == min(ny, t.y1 + 13)
        # columns: at least the halo (clipped to the image), at most 19 extra columns for alignment
        assert t.px0 <= max(0, t.x0 - 13) and t.px1 >= min(nx, t.x1 + 13)
        assert t.px0 >= max(0, t.x0 - 13 - 19) and t.px1 <= min(nx, t.x1 + 13 + 19)
        if nx % 4 == 0:
            assert (t.px1 - t.px0) % 4 == 0
        oy, ox = t.owned
        assert oy.stop - oy.start == t.y1 - t.y0 and ox.stop - ox.start == t.x1 - t.x0
    assert np.all(cover == 1)


def test_grid_shape_prefers_square_tiles():
    assert sorted(tiles.grid_shape(8, 320, 320)) == [2, 4]
    assert tiles.grid_shape(4, 320, 320) == (2, 2)
    assert tiles.grid_shape(2, 100, 400) == (1, 2)


def test_tile_index_mapping_round_trip():
    nz, ny, nx = 7, 40, 50
    plan = tiles.plan_tiles(ny, nx, 4, 3)
    full = np.arange(nz * ny * nx).reshape(nz, ny, nx)
    seen = []
    for t in plan:
        sub = full[(slice(None),) + t.padded]
        local = np.arange(sub.size)
        glob, keep = tiles.tile_linear_to_global(local, t, nz, ny, nx)
        assert np.array_equal(glob, sub.reshape(-1)[keep])
        assert np.all(np.diff(glob) > 0)                      # order preserved
        seen.append(glob)
    assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(nz * ny * nx))


class NumpyCounter:
    """Stand-in for the K4 kernels with the semantics of ogn_purity_stats / ogn_purity_counts."""

    def stats(self, ext, segmask):
        nz, ny, nx = ext.shape
        img = ny * nx
        sp = np.zeros(img, dtype=np.float32)
        np.maximum.at(sp, ext.max_index % img, ext.max_value)
        mv = ext.min_value if segmask is None else ext.min_value[~segmask.reshape(-1)[ext.min_index % img]]
        return (float(ext.max_value.max()) if len(ext.max_value) else -np.inf,
                float(mv.max()) if len(mv) else -np.inf, sp.reshape(ny, nx))

    def counts(self, ext, segmask, thr):
        nz, ny, nx = ext.shape
        img = ny * nx
        mv = ext.min_value if segmask is None else ext.min_value[~segmask.reshape(-1)[ext.min_index % img]]
        n1 = np.array([np.count_nonzero(ext.max_value.astype(np.float64) > t) for t in thr], dtype=np.int64)
        n0 = np.array([np.count_nonzero(mv.astype(np.float64) > t) for t in thr], dtype=np.int64)
        return n1, n0


def _fake_extrema(seed, shape):
    rng = np.random.default_rng(seed)
    lmax = np.where(rng.random(shape) < 0.03, rng.gamma(2.0, 2.0, shape), 0).astype(np.float32)
    lmin = np.where(rng.random(shape) < 0.03, rng.gamma(2.0, 1.6, shape), 0).astype(np.float32)
    seg = (rng.random(shape[1:]) < 0.25).astype(np.int16)
    return lmax, lmin, seg


def _worker(rank, world, port, shape, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from origin_b200 import distributed as ogd
        from origin_b200 import lib_origin
        lmax, lmin, seg = _fake_extrema(5, shape)
        nz, ny, nx = shape
        plan = tiles.plan_tiles(ny, nx, world, 2)
        t = plan[rank]
        # what a rank holds after step05 on its padded sub-cube: lists relative to the sub-cube
        sub_max = lmax[(slice(None),) + t.padded]
        sub_min = lmin[(slice(None),) + t.padded]
        mi = np.flatnonzero(sub_max)
        ni = np.flatnonzero(sub_min)
        ext = lib_origin.LocalExtrema(sub_max.shape, mi, sub_max.reshape(-1)[mi], ni, sub_min.reshape(-1)[ni])
        ext = ogd.owned_extrema(ext, t, shape)
        red = ogd.Reducer()
        thr, tab = lib_origin.Compute_threshold_purity(0.7, ext, None, seg, None, allreduce=red,
                                                       _backend=NumpyCounter())
        lsum = np.full(4, float(rank + 1))
        lcnt = np.ones(4)
        red.lambda_mean(lsum, lcnt)
        out[rank] = (thr, {k: np.array(v) for k, v in tab.items()}, lsum.copy(), lcnt.copy(), red.calls)
    finally:
        dist.destroy_process_group()


def test_purity_threshold_allreduce_world2():
    shape = (12, 30, 26)
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, shape, out), nprocs=2, join=True)
    lmax, lmin, seg = _fake_extrema(5, shape)
    rthr, rtab = orc.threshold_purity(0.7, lmax.astype(np.float64), lmin.astype(np.float64), seg)
    for rank in (0, 1):
        thr, tab, lsum, lcnt, calls = out[rank]
        assert thr == pytest.approx(rthr, rel=1e-12) or (np.isinf(thr) and np.isinf(rthr))
        np.testing.assert_allclose(tab['Tval_r'], rtab['Tval_r'], rtol=1e-12)
        np.testing.assert_array_equal(tab['Det_M'], rtab['Det_M'])
        np.testing.assert_array_equal(tab['Det_m'], rtab['Det_m'])
        np.testing.assert_array_equal(lsum, np.full(4, 3.0))
        np.testing.assert_array_equal(lcnt, np.full(4, 2.0))
        assert calls >= 4                               # counts, two maxima, image, histogram, lambda sums


def test_owned_extrema_torch_and_numpy_branches_agree():
    """``owned_extrema`` has a numpy branch (host lists) and a torch branch (device lists of the asynchronous steps):
    same owned subset, same global indices, same order, for every tile of a ragged plan."""
    import torch
    from origin_b200 import distributed as ogd, lib_origin
    nz, ny, nx = 7, 37, 53
    rng = np.random.default_rng(3)
    for t in tiles.plan_tiles(ny, nx, 6, 5):
        th, tw = t.shape
        n = nz * th * tw
        mi = np.sort(rng.choice(n, size=n // 9, replace=False)).astype(np.int64)
        ni = np.sort(rng.choice(n, size=n // 11, replace=False)).astype(np.int64)
        mv, nv = rng.normal(size=mi.size).astype(np.float32), rng.normal(size=ni.size).astype(np.float32)
        a = ogd.owned_extrema(lib_origin.LocalExtrema((nz, th, tw), mi, mv, ni, nv), t, (nz, ny, nx))
        b = ogd.owned_extrema(lib_origin.LocalExtrema((nz, th, tw), torch.from_numpy(mi), torch.from_numpy(mv),
                                                      torch.from_numpy(ni), torch.from_numpy(nv)), t, (nz, ny, nx))
        np.testing.assert_array_equal(a.max_index, b.max_index.numpy())
        np.testing.assert_array_equal(a.max_value, b.max_value.numpy())
        np.testing.assert_array_equal(a.min_index, b.min_index.numpy())
        np.testing.assert_array_equal(a.min_value, b.min_value.numpy())
        assert a.shape == b.shape == (nz, ny, nx) and len(a.max_index) > 0
        z, y, x = np.unravel_index(a.max_index, (nz, ny, nx))
        assert y.min() >= t.y0 and y.max() < t.y1 and x.min() >= t.x0 and x.max() < t.x1
        assert np.all(np.diff(a.max_index.reshape(-1)) != 0)
