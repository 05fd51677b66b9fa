"""Raw bandwidth of the peer gather (ogn_scatter_tile) into rank 0, no compute (development aid; torchrun)."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from origin_b200 import lib_origin, tiles
from origin_b200 import distributed as ogd

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=dev)
nz, ny, nx = 3681, 320, 320
ctx = lib_origin.default_context()
for grid in ('auto', 'rows'):
    if grid == 'auto':
        plan = tiles.plan_tiles(ny, nx, world, 13)
    else:   # full-width row bands: every owned region is one contiguous run per plane
        ys = np.linspace(0, ny, world + 1).round().astype(int)
        plan = [tiles.Tile(r, int(ys[r]), int(ys[r + 1]), 0, nx, max(0, int(ys[r]) - 13), min(ny, int(ys[r + 1]) + 13), 0, nx)
                for r in range(world)]
    t = plan[rank]
    src = torch.randn((nz,) + t.shape, device=dev)
    pg = ogd.PeerGather(ctx, (nz, ny, nx), dst=0, slots=1)
    for rep in range(3):
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        for _ in range(5):
            pg.scatter(src, t, (ny, nx), slot=0)
        pg.wait()
        dt = (time.perf_counter() - t0) / 5
        if rank == 0:
            gb = nz * ny * nx * 4 / 1e9
            print('%s grid=%s tile=%s: %.3f ms per gather, %.0f GB/s aggregate (%.0f GB/s into rank 0 over NVLink)'
                  % (os.environ.get('OGN_SCATTER_KERNEL') and 'kernel' or 'dma', grid, t.shape, dt * 1e3, gb / dt,
                     gb * (world - 1) / world / dt), flush=True)
    pg.close()
dist.destroy_process_group()
