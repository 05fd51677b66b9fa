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

    def sum_(self, tensor):
        """In-place SUM of a tensor that already lives on the group's device (no host round trip; the
        result is ordered after the call on the current stream, like any torch collective)."""
        self.dist.all_reduce(tensor, op=self.dist.ReduceOp.SUM, group=self.group)
        self.calls += 1
        return tensor

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


class _DeviceBuffer:
    """``__cuda_array_interface__`` view of a raw device allocation (lets torch wrap it without a copy)."""

    def __init__(self, ptr, shape, typestr='<f4'):
        self.__cuda_array_interface__ = dict(shape=tuple(int(s) for s in shape), typestr=typestr,
                                             data=(int(ptr), False), version=3, strides=None)


class PeerGather:
    """Gather of the owned tiles of a per-rank product into ``dst``'s full cube over NVLink peer memory
    (``ogn_peer_alloc`` / ``ogn_peer_open`` / ``ogn_scatter_tile``): no packing, no NCCL send/recv, no
    unpacking — every rank's copy kernel stores straight into the destination rank's buffer, on a side
    stream, so the transfer of one step overlaps the kernels of the next.

    ``slots`` destination cubes are allocated on ``dst`` (use 2 when consecutive steps overlap so that a
    step never overwrites the cube the previous one is still filling).  ``torch.distributed`` only
    carries the 64-byte IPC handles.  Raises :class:`origin_b200._lib.OgnError` when the GPUs cannot map
    each other's memory; callers may then fall back to :func:`gather_owned`.
    """

    def __init__(self, ctx, global_shape, dst=0, slots=1, group=None, dst_only=False):
        import ctypes
        import torch
        import torch.distributed as dist
        self.ctx, self.dst, self.group, self.dist, self.dst_only = ctx, dst, group, dist, dst_only
        self.rank = dist.get_rank(group)
        self.shape = tuple(int(s) for s in global_shape)
        self.nbytes = int(np.prod(self.shape)) * 4
        self.ptrs, self._owned, self._mapped = [], [], []
        self._torch, self.stagger_us = torch, 0      # close() reads these, also when the mapping below fails
        handles = []
        if self.rank == dst:
            for _ in range(slots):
                p = ctypes.c_void_p()
                h = ctypes.create_string_buffer(64)
                ctx.check(ctx.lib.ogn_peer_alloc(ctx.handle, self.nbytes, ctypes.byref(p), h))
                self._owned.append(p.value)
                handles.append(bytes(h.raw))
        box = [handles]
        dist.broadcast_object_list(box, src=dst, group=group)
        handles = box[0]
        err = None
        if self.rank == dst:
            self.ptrs = list(self._owned)
        else:
            try:
                for h in handles:
                    p = ctypes.c_void_p()
                    ctx.check(ctx.lib.ogn_peer_open(ctx.handle, ctypes.create_string_buffer(h, 64), ctypes.byref(p)))
                    self._mapped.append(p.value)
                self.ptrs = list(self._mapped)
            except Exception as exc:  # noqa: BLE001 - reported collectively below
                err = exc
        # all ranks agree on whether the mapping worked
        flags = [None] * dist.get_world_size(group)
        dist.all_gather_object(flags, err is None, group=group)
        if not all(flags):
            self.close()
            raise err if err is not None else RuntimeError('peer mapping failed on another rank')

    def stagger(self, tile_bytes, link_gbs=650.0, microseconds=None):
        """Let the source ranks take turns on the destination's NVLink ingress: rank r (counted among the sources)
        starts its copy r turns after it was enqueued, a turn being the time one tile needs on the link alone
        (``tile_bytes / link_gbs``, or ``microseconds``).  Call it on every rank; it relies on the copies being
        enqueued at about the same time everywhere (e.g. right after a collective)."""
        world = self.dist.get_world_size(self.group)
        turn = microseconds if microseconds is not None else tile_bytes / (link_gbs * 1e3)
        order = [r for r in range(world) if r != self.dst]
        delay = int(round(order.index(self.rank) * turn)) if self.rank != self.dst else 0
        self.stagger_us = delay
        self.ctx.check(self.ctx.lib.ogn_peer_set_delay(self.ctx.handle, delay))

    def attach(self, slot=0):
        """Call before ``step05(..., tile=...)``: the step then delivers the tile itself (``ogn_set_local_gather``)
        and :meth:`scatter` has nothing left to do — on the destination rank the spectral kernel stores the voxels
        that rank owns straight into slot ``slot``; on the other ranks the peer copy starts right behind the
        spectral kernel (``dst_only=True`` in the constructor restricts this to the destination rank)."""
        self._attached = None
        if self.rank == self.dst or not self.dst_only:
            self.ctx.check(self.ctx.lib.ogn_set_local_gather(self.ctx.handle, self.ptrs[slot]))
            self._attached = slot

    def scatter(self, cube_tile, tile, global_hw, slot=0):
        """Enqueue the copy of the owned window of ``cube_tile`` (``[nz][th][tw]`` float32 device tensor)
        into slot ``slot`` of the destination; returns immediately."""
        from ._lib import ptr
        if getattr(self, '_attached', None) == slot:
            self._attached = None      # already delivered by the step itself (fused stores / copy behind K2)
            return
        nz, th, tw = cube_tile.shape
        gny, gnx = global_hw
        desc = np.array([gny, gnx, tile.py0, tile.px0, tile.y0 - tile.py0, tile.y1 - tile.py0, tile.x0 - tile.px0,
                         tile.x1 - tile.px0], dtype=np.int32)
        self.ctx.check(self.ctx.lib.ogn_scatter_tile(self.ctx.handle, ptr(cube_tile), nz, th, tw, ptr(desc),
                                                     self.ptrs[slot]))

    def join(self):
        """The context's stream waits (on the device) for the copies enqueued so far."""
        self.ctx.check(self.ctx.lib.ogn_peer_join(self.ctx.handle))

    def wait(self):
        """Host waits for this rank's copies, then a barrier: afterwards ``result()`` is complete."""
        self.ctx.check(self.ctx.lib.ogn_peer_sync(self.ctx.handle))
        self.dist.barrier(group=self.group)

    def result(self, slot=0):
        """The assembled ``[nz][ny][nx]`` tensor on the destination rank (a view of the peer buffer), else None."""
        if self.rank != self.dst:
            return None
        dev = self._torch.device('cuda', self.ctx.device)
        return self._torch.as_tensor(_DeviceBuffer(self.ptrs[slot], self.shape), device=dev)

    def close(self):
        if self.stagger_us:
            self.ctx.lib.ogn_peer_set_delay(self.ctx.handle, 0)
        # remote copies into the destination's buffers may still be in flight: every rank first waits for its
        # own copies, then all ranks meet, and only then are mappings closed and the owned buffers freed
        if self._mapped or self._owned:
            try:
                self.wait()
            except Exception:  # noqa: BLE001 - tearing down after a failure: free what we can
                pass
        for p in self._mapped:
            self.ctx.lib.ogn_peer_close(self.ctx.handle, p)
        for p in self._owned:
            self.ctx.lib.ogn_peer_free(self.ctx.handle, p)
        self._mapped, self._owned, self.ptrs = [], [], []
