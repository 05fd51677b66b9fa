#!/usr/bin/env python
"""Benchmark of ORIGIN's detection hot path on B200: step05 TGLR (+ local extrema + step06
purity counts) in Gvoxel.profiles/s on a synthetic MUSE-shaped cube.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--shape NZ NY NX] [--dico 3FWHM|2_12]

Prints ONE JSON line (see DESIGN.md "Measurement" for every field).

* ``value``  device-resident: inputs already in HBM, CUDA events around K steps.  Steps are
             asynchronous (``step05(sync=False)``): nothing of a step returns to the host, the
             final synchronisation of the timed region waits for all of them.
* ``e2e``    the same step through the public host API (``lib_origin.step05`` ->
             ``ogn_step05``) with pinned HOST buffers: every step copies the cube and
             mask host->device and every product device->host inside the timed region.
* ``roofline`` dominant kernel (K1, per-lambda FSF correlation): algorithmic FLOPs /
             CUDA-event duration on the launching stream, against the FP32 FFMA peak
             measured on this device by ``tools/fma_peak`` in the same run; ``executed`` gives
             the FP32 issue slots the row-folded kernel really uses, ``traffic`` its DRAM bytes
             per launch from the committed ncu capture (``profiles/ncu_summary.json``).
* ``cpu_baseline`` the float64 oracle port of the reference algorithm
             (``oracle/origin_oracle.py``) timed on the host cores on a bounded spatial
             tile of the same workload (rank 0, N=1 only).
* ``--impl reference``: that CPU port alone, in the same JSON shape.

N > 1 (launched by torchrun): the cube is split into spatial tiles with >= (P//2 + 1)-pixel
halos, one per rank (strong scaling of the fixed cube); the step adds the NCCL allreduce
of the per-threshold purity counts (device-resident, in place) and the gather of the owned
correl tiles into rank 0's cube over NVLink peer memory (``ogn_scatter_tile`` on a side
stream; rank 0's own tile is stored by its spectral kernel).  Environment switches
``OGN_BENCH_NO_GATHER`` / ``OGN_BENCH_SKIP_LOCAL`` / ``OGN_BENCH_SYNC_STEP`` are diagnostics.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PSF_SIZE = 25
ASYNC_STEP = not os.environ.get('OGN_BENCH_SYNC_STEP')     # diagnostic switch: host-synchronous steps
SHAPE = (3681, 320, 320)
CPU_TILE = (64, 96)          # spatial sample of the workload for the CPU baseline


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--shape', type=int, nargs=3, default=list(SHAPE))
    ap.add_argument('--dico', default='3FWHM', choices=['3FWHM', '2_12'])
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--trace', action='store_true', help='host wall-clock per phase of a step (stderr)')
    return ap.parse_args()


def dictionary(name):
    from origin_b200 import dictionaries
    return dictionaries.dico_3fwhm()[0] if name == '3FWHM' else dictionaries.dico_fwhm_2_12()[0]


def sum_taps(profs, pcut=1e-8):
    from origin_b200.lib_origin import prepare_profiles
    return sum(len(p) for p in prepare_profiles(profs, pcut, True))


# ------------------------------------------------------------------------------------------
# CPU leg (cpu_baseline and --impl reference): the oracle port on a bounded tile
# ------------------------------------------------------------------------------------------

def cpu_step(cube, fsf, profs, mask, workers):
    from oracle import origin_oracle as orc
    out = orc.tglr_step(cube, fsf, None, profs, mask, 3, workers, 1e-8, True)
    orc.threshold_purity(0.9, out['cube_local_max'], out['cube_local_min'])
    return out


def cpu_inputs(nz, seed=0):
    from origin_b200 import synthetic
    ty, tx = CPU_TILE
    fsf = synthetic.moffat_fsf(nz)
    cube, _ = synthetic.faint_cube((nz, ty, tx), fsf, n_src=4, seed=seed)
    mask = synthetic.footprint_mask((nz, ty, tx), seed=seed)
    return cube, fsf, mask


def run_cpu(nz, profs, steps, warmup):
    cores = os.cpu_count() or 1
    cube, fsf, mask = cpu_inputs(nz)
    for _ in range(warmup):
        cpu_step(cube, fsf, profs, mask, cores)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(cube, fsf, profs, mask, cores)
    dt = (time.perf_counter() - t0) / steps
    units = cube.size * len(profs) / 1e9
    return dict(value=units / dt, seconds_per_step=dt, cores=cores,
                sample='%dx%dx%d tile of the cube, all %d profiles, float64, scipy.fft workers=%d'
                       % (cube.shape + (len(profs), cores)))


def main_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    profs = dictionary(args.dico)
    nz = args.shape[0]
    res = run_cpu(nz, profs, max(1, min(args.steps, 10)), min(args.warmup, 3))
    line = dict(
        metric='step05 TGLR throughput', value=res['value'], unit='Gvoxel.profiles/s', impl='reference',
        n_gpus=args.gpus, steps=max(1, min(args.steps, 10)), warmup=min(args.warmup, 3),
        ms_per_step=res['seconds_per_step'] * 1e3, higher_is_better=True, scaling='strong', vs_baseline=None,
        dtype='f64', data='synthetic',
        config=dict(workload='step05 TGLR + local extrema + step06 purity counts, %dx%dx%d cube, Dico_%s'
                             % (tuple(args.shape) + (args.dico,)), psf_size=PSF_SIZE),
        cpu_baseline=dict(value=res['value'], unit='Gvoxel.profiles/s', cores=res['cores'], kind='port',
                          sample=res['sample']),
        e2e=dict(value=res['value'], unit='Gvoxel.profiles/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
        note='oracle port of the reference algorithm (the Python reference cannot travel to the GPU box); '
             'each step is a bounded spatial tile of the workload',
    )
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# GPU leg
# ------------------------------------------------------------------------------------------

class ClockSampler:
    QUERY = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def fma_peak():
    exe = os.path.join(ROOT, 'tools', 'fma_peak')
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout.strip().splitlines()[-1]
        return json.loads(out)
    except Exception as exc:  # noqa: BLE001
        return dict(error=str(exc))


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return json.load(f)
    except OSError:
        return {}


def ncu_summary():
    try:
        with open(os.path.join(ROOT, 'profiles', 'ncu_summary.json')) as f:
            return json.load(f)
    except OSError:
        return {}


def main_gpu(args):
    import torch
    import torch.distributed as dist
    from origin_b200 import _lib, lib_origin, synthetic, tiles
    from origin_b200 import distributed as ogd

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the ported path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    dev = torch.device('cuda', local_rank)
    ctx = _lib.default_context(local_rank)

    nz, ny, nx = args.shape
    profs = dictionary(args.dico)
    nprof = len(profs)
    fsf_host = synthetic.moffat_fsf(nz, PSF_SIZE)
    fsf = torch.from_numpy(fsf_host).to(dev)
    halo = PSF_SIZE // 2 + 1
    all_tiles = tiles.plan_tiles(ny, nx, world, halo)
    tile = all_tiles[rank]

    # synthetic global cube, identical on every rank (same device generator seed), then the rank's tile
    gen = torch.Generator(device=dev).manual_seed(0)
    cube_g = torch.randn((nz, ny, nx), device=dev, dtype=torch.float32, generator=gen)
    yy = torch.arange(ny, device=dev, dtype=torch.float32)[:, None]
    xx = torch.arange(nx, device=dev, dtype=torch.float32)[None, :]
    cy, cx, th = (ny - 1) / 2, (nx - 1) / 2, float(np.deg2rad(3.0))
    u = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)
    v = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
    foot = (u.abs() > 0.49 * nx) | (v.abs() > 0.49 * ny)                      # ~5 % of the spaxels
    mask_g = (torch.rand((nz, ny, nx), device=dev, generator=gen) < 1e-3) | foot[None]
    n_src = max(1, int(round(200 * nz * ny * nx / (3681 * 320 * 320))))
    rng = np.random.default_rng(0)
    for _ in range(n_src):                                                        # injected line emitters
        z0, y0, x0 = int(rng.integers(20, nz - 20)), int(rng.integers(13, ny - 13)), int(rng.integers(13, nx - 13))
        sig = rng.uniform(2.0, 12.0) / 2.3548
        hw = int(np.ceil(4 * sig))
        zz = np.arange(max(0, z0 - hw), min(nz, z0 + hw + 1))
        line = np.exp(-0.5 * ((zz - z0) / sig) ** 2)
        spat = fsf_host[z0]
        amp = rng.uniform(5.0, 30.0) / np.sqrt((spat ** 2).sum() * (line ** 2).sum())
        patch = torch.from_numpy((amp * line[:, None, None] * spat[None]).astype(np.float32)).to(dev)
        cube_g[zz[0]:zz[-1] + 1, y0 - 12:y0 + 13, x0 - 12:x0 + 13] += patch
    py, px = tile.padded
    cube = cube_g[:, py, px].contiguous()
    mask = mask_g[:, py, px].to(torch.uint8).contiguous()
    del cube_g, mask_g
    tz, ty_, tx_ = cube.shape
    vol_tile = tz * ty_ * tx_
    cap = max(4096, vol_tile // 40)
    def alloc_out():
        return dict(correl=torch.empty_like(cube), correl_min=torch.empty_like(cube),
                    profile=torch.empty(cube.shape, dtype=torch.uint8, device=dev),
                    maxmap=torch.empty((ty_, tx_), device=dev), minmap=torch.empty((ty_, tx_), device=dev),
                    max_index=torch.empty(cap, dtype=torch.int64, device=dev),
                    max_value=torch.empty(cap, device=dev),
                    min_index=torch.empty(cap, dtype=torch.int64, device=dev),
                    min_value=torch.empty(cap, device=dev))

    # N > 1: two product sets, so that the peer copy of step i (side stream) overlaps the kernels of step i+1
    out_sets = [alloc_out() for _ in range(2 if world > 1 else 1)]
    reducer = ogd.Reducer() if world > 1 else None
    gather, gather_mode = None, None
    if world > 1:
        try:
            gather = ogd.PeerGather(ctx, (nz, ny, nx), dst=0, slots=2)
            gather_mode = 'owned correl tiles stored into rank 0 over NVLink peer memory (ogn_scatter_tile, CUDA IPC)'
        except Exception as exc:  # noqa: BLE001
            gather_mode = 'NCCL send/recv (peer mapping unavailable: %s)' % str(exc)[:120]
    thresholds = np.linspace(4.0, 12.0, 50)
    thr_dev = torch.from_numpy(thresholds).to(dev)
    counts_dev = torch.zeros(2 * len(thresholds), dtype=torch.int64, device=dev)
    state = {'i': 0}

    trace = {}

    def tick(name, t0):
        if args.trace:
            trace[name] = trace.get(name, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    def step():
        i = state['i']
        state['i'] = i + 1
        t0 = time.perf_counter()
        if gather is not None and not os.environ.get('OGN_BENCH_NO_GATHER'):
            gather.attach(slot=i % 2)      # rank 0's own tile is stored by K2 itself
        # sync=False: nothing of the step comes back to the host (the extremum list lengths stay on the device),
        # so the host runs ahead and the GPU never idles between steps; the final synchronisation of the timed
        # region waits for everything
        res = lib_origin.step05(cube, fsf, None, profs, mask, 3, 1e-8, True, out=out_sets[i % len(out_sets)], ctx=ctx,
                                tile=(tile, (ny, nx)) if world > 1 else None, sync=not ASYNC_STEP)
        t0 = tick('step05', t0)
        ext = res['extrema']                                                      # owned voxels, global indices
        # step06 counting loop on the device lists; the counts stay on the device ...
        n1, n0 = lib_origin.purity_counts(ext, None, thr_dev, ctx, out=counts_dev)
        t0 = tick('purity_counts', t0)
        if world > 1:
            reducer.sum_(counts_dev)                                              # ... NCCL allreduce of the histograms
            t0 = tick('allreduce', t0)
            # correl -> rank 0.  Enqueued last: the bulk stores would otherwise sit in front of the small
            # allreduce on the NVLink queues; this way they overlap the kernels of the next step instead.
            if os.environ.get('OGN_BENCH_NO_GATHER') or (os.environ.get('OGN_BENCH_SKIP_LOCAL') and rank == 0):
                pass                                      # diagnostics only: isolate the cost of the gather
            elif gather is not None:
                gather.scatter(res['correl'], tile, (ny, nx), slot=i % 2)
            else:
                state['correl_full'] = ogd.gather_owned(res['correl'], tile, all_tiles, (nz, ny, nx))
            t0 = tick('gather', t0)
        state['n1'], state['n0'], state['ext'] = n1, n0, ext
        return res

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync_all()
    ctx.timing(True)
    ctx.timing_report()
    trace.clear()
    launches0 = ctx.launch_count
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(args.steps):
        step()
    if gather is not None:
        gather.join()                     # the timed region ends after the last peer copy
    e1.record()
    sync_all()
    ms_total = e0.elapsed_time(e1)
    if args.trace:
        sys.stderr.write('rank %d host ms/step: %s\n' % (rank, ' '.join('%s=%.3f' % (k, v * 1e3 / args.steps)
                                                                          for k, v in trace.items())))
    clocks = sampler.stop()
    launches = ctx.launch_count - launches0
    state['folded'] = ctx.fsf_folded
    stages = {}
    for name, ms in ctx.timing_report():
        stages.setdefault(name, []).append(ms)
    ctx.timing(False)
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    ms_step = ms_total / args.steps
    units = nz * ny * nx * nprof / 1e9
    value = units / (ms_step * 1e-3)

    # ---- end to end through the host API with pinned host buffers --------------------------
    e2e = None
    if not args.no_e2e:
        cube_h = _lib.pinned_empty(cube.shape, np.float32)
        mask_h = _lib.pinned_empty(cube.shape, np.uint8)
        cube_h[...] = cube.cpu().numpy()
        mask_h[...] = mask.cpu().numpy()
        out_h = dict(correl=_lib.pinned_empty(cube.shape, np.float32), correl_min=_lib.pinned_empty(cube.shape, np.float32),
                     profile=_lib.pinned_empty(cube.shape, np.uint8), maxmap=_lib.pinned_empty((ty_, tx_), np.float32),
                     minmap=_lib.pinned_empty((ty_, tx_), np.float32),
                     max_index=_lib.pinned_empty((cap,), np.int64), max_value=_lib.pinned_empty((cap,), np.float32),
                     min_index=_lib.pinned_empty((cap,), np.int64), min_value=_lib.pinned_empty((cap,), np.float32))

        def step_e2e():
            res = lib_origin.step05(cube_h, fsf_host, None, profs, mask_h, 3, 1e-8, True, out=out_h, ctx=ctx)
            ext = res['extrema']
            n1, n0 = lib_origin.purity_counts(lib_origin.LocalExtrema(cube.shape, ext.max_index, ext.max_value,
                                                                       ext.min_index, ext.min_value), None,
                                              thresholds, ctx)
            if world > 1:
                reducer.sum(np.concatenate([n1, n0]))
            return res

        step_e2e()
        sync_all()
        n_e2e = max(2, min(args.steps, 5))
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n_e2e):
            res = step_e2e()
        e1.record()
        sync_all()
        wall = (time.perf_counter() - t0) * 1e3 / n_e2e
        ms_e2e = max(e0.elapsed_time(e1) / n_e2e, wall)
        if world > 1:
            t = torch.tensor([ms_e2e], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e2e = float(t.item())
        n1c, n0c = res['extrema'].counts
        h2d = cube_h.nbytes + mask_h.nbytes + fsf_host.nbytes
        d2h = (out_h['correl'].nbytes + out_h['correl_min'].nbytes + out_h['profile'].nbytes
               + out_h['maxmap'].nbytes + out_h['minmap'].nbytes + 12 * (n1c + n0c) + 16)
        e2e = dict(value=units / (ms_e2e * 1e-3), unit='Gvoxel.profiles/s', ms_per_step=ms_e2e,
                   h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                   api='origin_b200.lib_origin.step05 -> ogn_step05 (pinned host numpy in/out) + purity_counts')

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------
    peak = fma_peak() if world == 1 else {}
    mp = measured_peaks()
    k1_ms = float(np.mean(stages.get('k1_fsf_correlate', [np.nan])))
    k2_ms = float(np.mean(stages.get('k2_spectral_glr', [np.nan])))
    k3_ms = float(np.mean(stages.get('k3_local_extrema', [np.nan])))
    k1_flops = 2.0 * PSF_SIZE * PSF_SIZE * vol_tile
    k2_flops = 2.0 * sum_taps(profs) * vol_tile
    fp32_peak = peak.get('fp32_tflops')
    summ = ncu_summary()
    folded = bool(state.get('folded'))
    half = PSF_SIZE // 2
    # FP32-pipe instructions K1 executes per voxel (x2 = flop slots of the FFMA peak): direct form P^2 FFMA;
    # folded form (half+1) rows x P FFMA + half rows x (32+P-1)/32 FADD (one add per loaded input sample)
    k1_slots = 2.0 * ((half + 1) * PSF_SIZE + half * (32 + PSF_SIZE - 1) / 32.0) if folded else 2.0 * PSF_SIZE ** 2
    k1_alg = k1_flops / (k1_ms * 1e-3) / 1e12
    k1_exec = k1_slots * vol_tile / (k1_ms * 1e-3) / 1e12
    roofline = dict(
        bound='fp32', kernel='k1::fsf_correlate_kernel<25>' + (' (row-folded: mirror-symmetric FSF)' if folded else ''),
        achieved=k1_alg, peak=fp32_peak, unit='TFLOP/s', frac=(k1_alg / fp32_peak) if fp32_peak else None,
        executed=dict(flop_slots_per_voxel=k1_slots, tflops=k1_exec, frac=(k1_exec / fp32_peak) if fp32_peak else None,
                      note='FP32-pipe issue slots actually used (FFMA and FADD both count 2): the figure to read as '
                           'pipe utilisation; "achieved" counts the direct-form 2*P^2 flops of SURVEY.md 8d, which the '
                           'folded kernel does not execute, so it can exceed the peak'),
        frac_above_one=('K1 folds the mirror-symmetric FSF rows: it executes %.0f of the %d algorithmic flop slots per '
                        'voxel, so achieved/peak can exceed 1; executed.frac is the FP32 pipe utilisation'
                        % (k1_slots, 2 * PSF_SIZE ** 2)) if folded else None,
        traffic=summ.get('k1_dram_bytes_per_launch'),
        peak_source='FP32 FFMA peak measured on this device in this run by tools/fma_peak '
                    '(MEASURED_PEAKS.json has no FP32 figure; nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4)',
        algorithmic='2*P^2 = %d flop/voxel x %d voxels per launch' % (2 * PSF_SIZE ** 2, vol_tile),
        ms_per_launch=k1_ms,
        note='the schema value "tensor" does not apply: the path is FP32-FMA bound (SURVEY.md 8d), '
             'tensor cores cannot meet the 1e-5 parity bound',
    )
    hbm_peak = mp.get('hbm_gbs')
    kernels = dict(
        k1_fsf_correlate=dict(ms=k1_ms, tflops=k1_flops / (k1_ms * 1e-3) / 1e12, bound='fp32'),
        k2_spectral_glr=dict(ms=k2_ms, tflops=k2_flops / (k2_ms * 1e-3) / 1e12, bound='fp32'),
        k3_local_extrema=dict(ms=k3_ms, gbs=9.0 * vol_tile / (k3_ms * 1e-3) / 1e9, bound='hbm',
                              frac_of_measured_hbm=(9.0 * vol_tile / (k3_ms * 1e-3) / 1e9 / hbm_peak) if hbm_peak else None),
        other_ms={k: float(np.mean(v)) for k, v in stages.items()
                  if k not in ('k1_fsf_correlate', 'k2_spectral_glr', 'k3_local_extrema', 'step05_span')},
        step05_span_ms=float(np.mean(stages.get('step05_span', [np.nan]))),
    )
    total_flops = (2.0 * PSF_SIZE ** 2 + 2.0 * sum_taps(profs)) * nz * ny * nx
    line = dict(
        metric='step05 TGLR throughput', value=value, unit='Gvoxel.profiles/s', n_gpus=world, steps=args.steps,
        warmup=args.warmup, ms_per_step=ms_step, higher_is_better=True, scaling='strong', vs_baseline=None,
        dtype='f32', data='synthetic',
        config=dict(workload='step05 TGLR + local extrema + step06 purity counts, %dx%dx%d float32 cube, Dico_%s '
                             '(%d profiles), single field, mask ~5%% footprint + 0.1%% voxels'
                             % (nz, ny, nx, args.dico, nprof),
                    psf_size=PSF_SIZE, parallelism='spatial tiles %s with %d-px halos' % (
                        'x'.join(str(v) for v in tiles.grid_shape(world, ny, nx)), halo),
                    gather=gather_mode,
                    l2='inputs (%.1f GB per rank) exceed the 126 MB L2; no flush needed' % (cube.numel() * 4 / 1e9),
                    timed_region='CUDA events on the launching stream around %d steps, barrier + synchronize on both '
                                 'sides, max over ranks' % args.steps),
        e2e=e2e, gpu_launches=int(launches), clocks=clocks, roofline=roofline, kernels=kernels,
        step_fp32_tflops=total_flops / (ms_step * 1e-3) / 1e12,
        step_fp32_frac=(total_flops / (ms_step * 1e-3) / 1e12 / fp32_peak) if fp32_peak else None,
        fma_peak=peak, n_local_extrema=[int(v) for v in state['ext'].counts],
    )
    if world == 1 and not args.no_cpu:
        cb = run_cpu(nz, profs, 1, 0)
        line['cpu_baseline'] = dict(value=cb['value'], unit='Gvoxel.profiles/s', cores=cb['cores'], kind='port',
                                    sample=cb['sample'], seconds_for_sample=cb['seconds_per_step'])
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse_args()
    if a.impl == 'reference':
        main_reference(a)
    else:
        main_gpu(a)
