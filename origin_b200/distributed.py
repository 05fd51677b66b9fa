"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over
NVLink 5 / NVSwitch on the B200 box, gloo in the CPU tests).

The path shards by independent spatial tiles (:mod:`origin_b200.tiles`); the
only exchange steps are tiny reductions — the per-threshold purity counts
(int64 SUM, reference lib_origin.py:1443-1449), two scalars and one image
(MAX, :1437-1438), the per-wavelength sums of step01 (float64 SUM,
steps.py:442) — and the gather of the owned ``correl`` tiles to rank 0.
"""

import numpy as np

from . import tiles as _tiles


class Reducer:
    """Numpy-facing reductions over a ``torch.distributed`` process group."""

    def __init__(self, group=None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if device is None:
            device = 'cuda' if dist.get_backend(group) == 'nccl' else 'cpu'
        self.device = device
        self.calls = 0

    def _reduce(self, arr, op, dtype):
        t = self.torch.as_tensor(np.ascontiguousarray(arr, dtype=dtype)).to(self.device)
        self.dist.all_reduce(t, op=op, group=self.group)
        self.calls += 1
        return t.cpu().numpy()

    def sum(self, arr):
        """Element-wise SUM of an int64 (or float64) array across ranks."""
        arr = np.asarray(arr)
        dtype = np.int64 if arr.dtype.kind in 'iu' else np.float64
        return self._reduce(arr, self.dist.ReduceOp.SUM, dtype)

    def max(self, arr):
        return self._reduce(arr, self.dist.ReduceOp.MAX, np.float64)

    def max_image(self, img):
        return self._reduce(img, self.dist.ReduceOp.MAX, np.float32)

    def lambda_mean(self, lsum, lcnt):
        """In-place SUM of the step01 per-wavelength partial sums / counts."""
        both = self.sum(np.concatenate([lsum, lcnt]).astype(np.float64))
        lsum[:] = both[:len(lsum)]
        lcnt[:] = both[len(lsum):]

    def barrier(self):
        self.dist.barrier(group=self.group)


def owned_extrema(ext, tile, global_shape):
    """Restrict a tile's :class:`~origin_b200.lib_origin.LocalExtrema` (indices
    relative to the padded sub-cube) to the voxels the tile owns and re-express
    them as global linear indices; order (C order within the tile) is kept."""
    from .lib_origin import LocalExtrema, _is_torch
    nz, ny, nx = global_shape

    def conv(index, value):
        if _is_torch(index):
            th, tw = tile.shape
            z = index // (th * tw)
            rem = index - z * (th * tw)
            y = rem // tw + tile.py0
            x = rem - (rem // tw) * tw + tile.px0
            keep = (y >= tile.y0) & (y < tile.y1) & (x >= tile.x0) & (x < tile.x1)
            glob = (z * ny + y) * nx + x
            return glob[keep], value[keep]
        glob, keep = _tiles.tile_linear_to_global(index, tile, nz, ny, nx)
        return glob, np.asarray(value)[keep]

    mi, mv = conv(ext.max_index, ext.max_value)
    ni, nv = conv(ext.min_index, ext.min_value)
    return LocalExtrema(global_shape, mi, mv, ni, nv)


def gather_owned(cube_tile, tile, all_tiles, global_shape, dst=0, group=None):
    """Gather the owned part of a per-rank product cube ``[nz][th][tw]`` (torch
    tensor) to rank ``dst``; returns the assembled ``[nz][ny][nx]`` tensor there
    and None elsewhere.  Point-to-point sends (the tiles differ in shape)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    ys, xs = tile.owned
    mine = cube_tile[:, ys, xs].contiguous()
    if rank != dst:
        dist.send(mine, dst=dst, group=group)
        return None
    nz, ny, nx = global_shape
    out = torch.empty((nz, ny, nx), dtype=cube_tile.dtype, device=cube_tile.device)
    for t in all_tiles:
        gy, gx = t.global_owned
        if t.rank == dst:
            out[:, gy, gx] = mine
        else:
            buf = torch.empty((nz, t.y1 - t.y0, t.x1 - t.x0), dtype=cube_tile.dtype, device=cube_tile.device)
            dist.recv(buf, src=t.rank, group=group)
            out[:, gy, gx] = buf
    return out
