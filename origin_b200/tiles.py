"""Spatial tile planner for multi-GPU runs.

Every stage of the hot path is local in (y, x) — radius ``P//2`` for the FSF
correlation plus 1 for the 3x3x3 extremum — and needs the full wavelength axis,
so the cube is partitioned into a grid of spatial tiles, one per rank, each
carrying a halo cut from the *input* cube (no inter-GPU halo exchange).  A rank
runs the unchanged single-GPU kernels on its padded sub-cube; a tile computed
with a halo reproduces the full-cube result bit for bit in its interior
(``tests/test_gpu_parity.py::test_tglr_tile_consistency_and_spot_oracle``).
"""

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Tile:
    rank: int
    y0: int          # owned region [y0, y1) x [x0, x1) in global coordinates
    y1: int
    x0: int
    x1: int
    py0: int         # padded region (owned + halo, clipped to the image)
    py1: int
    px0: int
    px1: int

    @property
    def owned(self):
        """Slices of the owned region inside the padded sub-cube."""
        return (slice(self.y0 - self.py0, self.y1 - self.py0), slice(self.x0 - self.px0, self.x1 - self.px0))

    @property
    def padded(self):
        return (slice(self.py0, self.py1), slice(self.px0, self.px1))

    @property
    def global_owned(self):
        return (slice(self.y0, self.y1), slice(self.x0, self.x1))

    @property
    def shape(self):
        return (self.py1 - self.py0, self.px1 - self.px0)


def grid_shape(n, ny, nx):
    """Rows x columns of the tile grid for ``n`` ranks: the factorisation whose
    tiles are closest to square (1x2, 2x2, 2x4 for 2/4/8 on a square field)."""
    best = None
    for gy in range(n, 0, -1):               # ties go to more rows: wide tiles keep the 32-lane x groups full
        if n % gy:
            continue
        gx = n // gy
        th, tw = ny / gy, nx / gx
        score = (th + tw) / (th * tw)            # halo perimeter per unit area
        if best is None or score < best[0] - 1e-12:
            best = (score, gy, gx)
    return best[1], best[2]


def _padded_columns(x0, x1, nx, halo):
    """Column range ``[px0, px1)`` of a tile's sub-cube: at least ``halo`` columns either side of the
    owned ``[x0, x1)`` (clipped to the image), widened by a few columns so that the kernels find their
    fast paths — a sub-cube width that is a multiple of 16 (at least of 4: TMA tensor maps, float4 /
    uchar4 accesses and 16-byte mask rows need it; otherwise the library stages padded copies and falls
    back to scalar kernels) and, when possible, a computed window (owned minus one ring column) that
    starts on a 16-column boundary of the sub-cube.  Extra halo columns are real data and change no
    result."""
    best = None
    for a in range(20):
        p0 = max(0, x0 - halo - a)
        for b in range(20):
            p1 = min(nx, x1 + halo + b)
            w = p1 - p0
            wx0 = max(0, x0 - 1 - p0)                     # first column of the computed window
            score = (4 * (w % 16 == 0) + 2 * (w % 4 == 0) + (wx0 % 16 == 0), -(w))
            if best is None or score > best[0]:
                best = (score, p0, p1)
    return best[1], best[2]


def plan_tiles(ny, nx, n, halo):
    """``n`` tiles covering a (ny, nx) field, each padded by at least ``halo`` pixels where the image
    continues (columns a little more, see :func:`_padded_columns`)."""
    gy, gx = grid_shape(n, ny, nx)
    ys = np.linspace(0, ny, gy + 1).round().astype(int)
    xs = np.linspace(0, nx, gx + 1).round().astype(int)
    tiles = []
    for r in range(n):
        iy, ix = divmod(r, gx)
        y0, y1, x0, x1 = int(ys[iy]), int(ys[iy + 1]), int(xs[ix]), int(xs[ix + 1])
        px0, px1 = _padded_columns(x0, x1, nx, halo)
        tiles.append(Tile(r, y0, y1, x0, x1, max(0, y0 - halo), min(ny, y1 + halo), px0, px1))
    return tiles


def tile_linear_to_global(index, tile, nz, ny, nx):
    """Map C-order linear voxel indices of a tile's padded sub-cube to global
    linear indices, keeping only voxels of the owned region.  Returns
    ``(global_index, keep_mask)``; order is preserved within the tile."""
    th, tw = tile.shape
    index = np.asarray(index, dtype=np.int64)
    z, rem = np.divmod(index, th * tw)
    y, x = np.divmod(rem, tw)
    gy, gx = y + tile.py0, x + tile.px0
    keep = (gy >= tile.y0) & (gy < tile.y1) & (gx >= tile.x0) & (gx < tile.x1)
    glob = (z * ny + gy) * nx + gx
    return glob[keep], keep
