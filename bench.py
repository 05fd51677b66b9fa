#!/usr/bin/env python
"""Benchmark of ORIGIN's detection hot path on B200: step05 TGLR (+ local extrema + step06
purity counts) in Gvoxel.profiles/s on a synthetic MUSE-shaped cube.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--shape NZ NY NX] [--dico 3FWHM|2_12]

Prints ONE JSON line (DESIGN.md "Measurement" explains every field).

* ``value``    device-resident: inputs already in HBM, CUDA events around K asynchronous steps
               (``step05(sync=False)``: nothing of a step returns to the host).
* ``e2e``      the same step through the public host API (``lib_origin.step05`` -> ``ogn_step05``)
               with pinned HOST buffers; host<->device copies inside the timed region.
* ``roofline`` the DOMINANT kernel of the configuration that ran (largest CUDA-event stage time):
               ``frac`` = FP32 issue slots the kernel executes / measured FFMA peak (``tools/fma_peak``
               in the same run), ``achieved_algorithmic`` = SURVEY 8d's algorithmic flops / time;
               HBM-bound kernels against ``MEASURED_PEAKS.json``.  ``kernels`` lists every stage.
* ``parity_spot``  after the timed region: the float64 direct-space oracle evaluated on >= 64 seeded
               voxels of the benchmark cube itself (corners, edges, tile seams, detections) and
               on their 3x3x3 neighbourhoods; worst error over the 1e-5 bound, argmax / extremum
               ties (north_star: "the number of such ties is reported").
* ``sharded_parity`` (N > 1) rank 0 recomputes the whole cube on its own GPU and compares with the
               peer-gathered cube, the merged extremum lists and the all-reduced counts.
* ``configs``  the other north_star configurations, driver-timed in the same run: ``c3`` =
               Dico_FWHM_2_12 + step01 (N = 1), ``c5`` = 3681x900x900 Dico_FWHM_2_12 on all ranks
               plus its 1-GPU time (N = 8).
* ``cpu_baseline`` / ``--impl reference``  the reference's own functions (``oracle/_ref``, the
               unmodified ``lib_origin.py``; ``kind: "reference"``) on the host cores, on tiles of
               the SAME cube; falls back to the oracle port (``kind: "port"``) without ``oracle/_ref``.

N > 1 (torchrun): the cube is split into spatial tiles with >= (P//2 + 1)-pixel halos, one per rank
(strong scaling of the fixed cube); a step adds the NCCL allreduce of the per-threshold purity counts
and the gather of the owned correl tiles into rank 0's cube over NVLink peer memory.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

if '--impl' in sys.argv and 'reference' in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1, which would throttle the CPU arm at N > 1 only
    os.environ.pop('OMP_NUM_THREADS', None)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PSF_SIZE = 25
ASYNC_STEP = not os.environ.get('OGN_BENCH_SYNC_STEP')     # diagnostic switch: host-synchronous steps
SHAPE = (3681, 320, 320)
C5_SHAPE = (3681, 900, 900)
THRESHOLDS = np.linspace(4.0, 12.0, 50)
RTOL = 1e-5                  # north_star parity bound: |d| <= RTOL * max(|ref|, rms(ref))


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
    ap.add_argument('--no-parity', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='skip the extra north_star configurations (c3 / c5)')
    ap.add_argument('--trace', action='store_true', help='host wall-clock per phase of a step (stderr)')
    return ap.parse_args()


def dictionary(name):
    from origin_b200 import dictionaries
    return dictionaries.dico_3fwhm()[0] if name == '3FWHM' else dictionaries.dico_fwhm_2_12()[0]


def cut_profiles(profs, pcut=1e-8):
    from origin_b200.lib_origin import prepare_profiles
    return prepare_profiles(profs, pcut, True)


def workload_config(shape, dico, nprof, n_gpus):
    """The ``config`` both arms print (identical: the driver compares them)."""
    from origin_b200 import tiles
    nz, ny, nx = shape
    return dict(workload='step05 TGLR + local extrema + step06 purity counts, %dx%dx%d float32 cube, Dico_%s '
                         '(%d profiles), single field, mask ~5%% footprint + 0.1%% voxels'
                         % (nz, ny, nx, dico, nprof), psf_size=PSF_SIZE,
                parallelism='spatial tiles %s with %d-px halos' % (
                    'x'.join(str(v) for v in tiles.grid_shape(max(1, n_gpus), ny, nx)), PSF_SIZE // 2 + 1),
                l2='inputs (%.1f GB for the cube, >= %.2f GB per rank) exceed the 126 MB L2; no flush needed'
                   % (nz * ny * nx * 4 / 1e9, nz * ny * nx * 4 / 1e9 / max(1, n_gpus)))


# ------------------------------------------------------------------------------------------
# CPU leg (cpu_baseline and --impl reference): the reference's own functions on tiles of the cube
# ------------------------------------------------------------------------------------------

_CPU = {}
_CPU_TILES = {}


def _cpu_tile_job(job):
    """One tile through the reference's step05 + step06 array path (runs in a forked worker)."""
    kind, origin, tile, gshape, dico = job
    from origin_b200 import synthetic
    nz = gshape[0]
    fsf = _CPU['fsf']
    profs = _CPU['profs']
    (y0, x0), (ty, tx) = origin, tile
    key = (origin, tile, tuple(gshape))
    if key not in _CPU_TILES:         # inputs are made outside the timed part and kept by the worker
        cube, mask = synthetic.bench_window(gshape, ((0, nz), (y0, y0 + ty), (x0, x0 + tx)), fsf)
        _CPU_TILES[key] = (cube, mask.astype(bool))
    cube, mask = _CPU_TILES[key]
    t0 = time.perf_counter()
    if kind == 'reference':
        lib = _CPU['lib']
        correl, profile, cmin = lib.Correlation_GLR_test(cube, fsf, None, profs, nthreads=1, pcut=1e-8, pmeansub=True)
        correl[mask] = 0                               # steps.py:781
        profile[mask] = 0                              # steps.py:788
        lmax, lmin = lib.compute_local_max(correl, cmin, mask, 3)
        lib.Compute_threshold_purity(0.9, lmax, lmin, threshlist=THRESHOLDS)
    else:
        from oracle import origin_oracle as orc
        out = orc.tglr_step(cube, fsf, None, profs, mask, 3, 1, 1e-8, True)
        orc.threshold_purity(0.9, out['cube_local_max'], out['cube_local_min'], threshlist=THRESHOLDS)
    return time.perf_counter() - t0


class CpuArm:
    """The reference's CPU implementation on the host cores: one process per core, each running the
    unmodified single-threaded reference functions on its own spatial tile of the benchmark cube (the
    reference's joblib threads do not scale: 1.0-1.4x on 8 threads, GIL-bound; independent tiles are how its
    path uses a whole host).  A step = ``cores`` tiles; the tile size is calibrated so that a step lasts
    ``target_s`` seconds."""

    def __init__(self, gshape, dico, target_s):
        import multiprocessing as mp
        from origin_b200 import synthetic
        from oracle import ref_loader
        self.gshape, self.dico = tuple(gshape), dico
        try:
            self.cores = len(os.sched_getaffinity(0))
        except AttributeError:
            self.cores = os.cpu_count() or 1
        _CPU['fsf'] = synthetic.moffat_fsf(gshape[0], PSF_SIZE)
        _CPU['profs'] = dictionary(dico)
        self.kind = 'port'
        if ref_loader.available():
            import warnings
            warnings.filterwarnings('ignore')
            _CPU['lib'] = ref_loader.load_lib_origin()
            self.kind = 'reference'
        self.nprof = len(_CPU['profs'])
        self.pool = mp.get_context('fork').Pool(self.cores)
        # calibration: one small tile per core
        ny, nx = gshape[1], gshape[2]
        cal = (min(16, ny), min(16, nx))
        t = self._run(cal)
        per_voxel = t / (gshape[0] * cal[0] * cal[1])
        want = max(cal[0] * cal[1], target_s / per_voxel / gshape[0])
        ty = int(min(ny // 2 if ny >= 64 else ny, max(16, round(np.sqrt(want) / 8) * 8)))
        tx = int(min(nx // 2 if nx >= 64 else nx, max(16, round(want / ty / 8) * 8)))
        self.tile = (ty, tx)

    def _jobs(self, tile):
        ny, nx = self.gshape[1], self.gshape[2]
        ty, tx = tile
        rng = np.random.default_rng(12345)
        jobs = []
        for _ in range(self.cores):
            y0 = int(rng.integers(0, max(1, ny - ty + 1)))
            x0 = int(rng.integers(0, max(1, nx - tx + 1)))
            jobs.append((self.kind, (y0, x0), tile, self.gshape, self.dico))
        return jobs

    def _run(self, tile):
        # the tiles run concurrently, one per core: the step lasts as long as the slowest of them
        return max(self.pool.map(_cpu_tile_job, self._jobs(tile), chunksize=1))

    def step(self):
        return self._run(self.tile)

    @property
    def units_per_step(self):
        return self.gshape[0] * self.tile[0] * self.tile[1] * self.cores * self.nprof / 1e9

    def sample(self):
        what = ('unmodified reference functions (oracle/_ref: Correlation_GLR_test nthreads=1 + compute_local_max + '
                'Compute_threshold_purity)' if self.kind == 'reference' else
                'oracle port of the reference algorithm (oracle/_ref absent)')
        return ('%d tiles of %dx%dx%d cut from the %dx%dx%d benchmark cube per step, one process per core, float64, %s'
                % ((self.cores, self.gshape[0]) + self.tile + self.gshape + (what,)))

    def close(self):
        self.pool.close()
        self.pool.join()


def main_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    # the whole run (calibration + warm-up + steps) is bounded to a few minutes
    target = float(np.clip(150.0 / (steps + warmup), 1.0, 10.0))
    arm = CpuArm(args.shape, args.dico, target)
    for _ in range(warmup):
        arm.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        arm.step()
    dt = (time.perf_counter() - t0) / steps
    value = arm.units_per_step / dt
    line = dict(
        metric='step05 TGLR throughput', value=value, unit='Gvoxel.profiles/s', impl='reference',
        n_gpus=args.gpus, steps=steps, warmup=warmup, ms_per_step=dt * 1e3, higher_is_better=True, scaling='strong',
        vs_baseline=None, dtype='f64', data='synthetic', config=workload_config(args.shape, args.dico, arm.nprof, args.gpus),
        cpu_baseline=dict(value=value, unit='Gvoxel.profiles/s', cores=arm.cores, kind=arm.kind, sample=arm.sample()),
        e2e=dict(value=value, unit='Gvoxel.profiles/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
        note='each step is a bounded sample of the workload (tiles of the same seeded cube the GPU arm runs); '
             'value = voxel.profiles of the sample / wall time',
    )
    arm.close()
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# GPU leg: helpers
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
        # the median under load: idle samples (before / after the timed region) sit at the idle clock
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def fma_peak():
    exe = os.path.join(ROOT, 'tools', 'fma_peak')
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout.strip().splitlines()[-1]
        return json.loads(out)
    except Exception as exc:  # noqa: BLE001
        return dict(error=str(exc))


def _load_json(*parts):
    try:
        with open(os.path.join(ROOT, *parts)) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


def k2f_layout(prof_cut, gmax=10, zb=8):
    """FP32 instructions per voxel of the folded spectral kernel (mirrors ``ogn_k2f_prepare``): groups of
    ``gmax`` profiles, half-lengths padded to multiples of ``zb``; per group (H + 1) x slots FFMA/FMUL and H FADD."""
    halves = [(len(p) - 1) // 2 for p in prof_cut]
    fma = add = 0
    for g0 in range(0, len(halves), gmax):
        grp = halves[g0:g0 + gmax]
        hpad = max(zb, -(-grp[-1] // zb) * zb)
        fma += (hpad + 1) * gmax
        add += hpad
    return fma, add


def kernel_models(prof_cut, folded_k1, uses_k2f):
    """Per-voxel work of each kernel: algorithmic flops and the FP32 issue slots (x2 = flop
    slots of the FFMA peak) the shipped kernel executes; HBM bytes for the extremum pass."""
    half = PSF_SIZE // 2
    k1_slots = (2.0 * ((half + 1) * PSF_SIZE + half * (32 + PSF_SIZE - 1) / 32.0) if folded_k1
                else 2.0 * PSF_SIZE ** 2)
    sum_l = sum(len(p) for p in prof_cut)
    if uses_k2f:
        fma, add = k2f_layout(prof_cut)
        k2_slots = 2.0 * (fma + add + len(prof_cut))          # + the normalisation multiply per profile
    else:
        k2_slots = 2.0 * (sum(-(-len(p) // 4) * 4 for p in prof_cut) + len(prof_cut))   # taps padded to 4, + normalisation
    return dict(
        k1_fsf_correlate=dict(bound='fp32', algorithmic=2.0 * PSF_SIZE ** 2, executed=k1_slots,
                              kernel='k1::fsf_correlate_kernel<25>' + (' (row-folded: mirror-symmetric FSF)'
                                                                       if folded_k1 else '')),
        k2_spectral_glr=dict(bound='fp32', algorithmic=2.0 * sum_l, executed=k2_slots,
                             kernel='k2f::folded_glr_kernel' if uses_k2f else 'k2::spectral_glr_kernel'),
        k3_local_extrema=dict(bound='hbm', bytes=9.0, kernel='local_extrema3_tma_kernel'),
    )


def stage_rooflines(stages, models, vol, fp32_peak, hbm_peak, traffic):
    """Roofline entry per timed stage + the dominant one."""
    out = {}
    for name, m in models.items():
        ms = float(np.mean(stages[name])) if name in stages else None
        if not ms:
            continue
        e = dict(kernel=m['kernel'], bound=m['bound'], ms_per_launch=ms)
        if m['bound'] == 'fp32':
            alg = m['algorithmic'] * vol / (ms * 1e-3) / 1e12
            exe = m['executed'] * vol / (ms * 1e-3) / 1e12
            e.update(achieved=exe, achieved_algorithmic=alg, peak=fp32_peak, unit='TFLOP/s',
                     frac=exe / fp32_peak if fp32_peak else None,
                     frac_algorithmic=alg / fp32_peak if fp32_peak else None,
                     flop_per_voxel=dict(algorithmic=m['algorithmic'], executed=m['executed']))
        else:
            gbs = m['bytes'] * vol / (ms * 1e-3) / 1e9
            e.update(achieved=gbs, peak=hbm_peak, unit='GB/s', frac=gbs / hbm_peak if hbm_peak else None,
                     bytes_per_voxel=m['bytes'])
        e['traffic'] = traffic.get(name)
        out[name] = e
    dominant = max(out, key=lambda k: out[k]['ms_per_launch']) if out else None
    return out, dominant


# ------------------------------------------------------------------------------------------
# GPU leg: one configuration on the current process group
# ------------------------------------------------------------------------------------------

class Job:
    """step05 + step06 counts of one (shape, dictionary) on ``world`` ranks (this rank's tile)."""

    def __init__(self, env, shape, dico, participate=True, world=None):
        import torch
        from origin_b200 import lib_origin, synthetic, tiles
        from origin_b200 import distributed as ogd
        self.env, self.torch = env, torch
        self.shape, self.dico = tuple(shape), dico
        self.world = env.world if world is None else world
        self.rank = env.rank if self.world > 1 else 0
        nz, ny, nx = self.shape
        self.profs = dictionary(dico)
        self.prof_cut = cut_profiles(self.profs)
        self.nprof = len(self.profs)
        self.fsf_host = synthetic.moffat_fsf(nz, PSF_SIZE)
        self.fsf = torch.from_numpy(self.fsf_host).to(env.dev)
        self.halo = PSF_SIZE // 2 + 1
        self.all_tiles = tiles.plan_tiles(ny, nx, self.world, self.halo)
        self.tile = self.all_tiles[self.rank]
        t = self.tile
        self.cube, self.mask = synthetic.bench_window(self.shape, ((0, nz), (t.py0, t.py1), (t.px0, t.px1)),
                                                      self.fsf_host, seed=0, xp=torch, device=env.dev)
        tz, ty, tx = self.cube.shape
        self.vol_tile = tz * ty * tx
        self.cap = max(4096, self.vol_tile // 24)
        dev = env.dev

        def alloc_out():
            return dict(correl=torch.empty_like(self.cube), correl_min=torch.empty_like(self.cube),
                        profile=torch.empty(self.cube.shape, dtype=torch.uint8, device=dev),
                        maxmap=torch.empty((ty, tx), device=dev), minmap=torch.empty((ty, tx), device=dev),
                        max_index=torch.empty(self.cap, dtype=torch.int64, device=dev),
                        max_value=torch.empty(self.cap, device=dev),
                        min_index=torch.empty(self.cap, dtype=torch.int64, device=dev),
                        min_value=torch.empty(self.cap, device=dev))

        # N > 1: three product sets, so that the peer copy of step i (side stream, staggered by rank) has until
        # the spectral kernel of step i+3 to finish
        self.out_sets = [alloc_out() for _ in range(3 if self.world > 1 else 1)]
        self.reducer = ogd.Reducer() if self.world > 1 else None
        self.gather, self.gather_mode = None, None
        if self.world > 1:
            try:
                # the peer copies are enqueued after the allreduce, which aligns the ranks: they then take turns on rank
                # 0's NVLink ingress (OGN_BENCH_GATHER_EARLY=1: copies start right behind K2, all at once;
                # OGN_BENCH_STAGGER_US: length of a turn, 0 = off)
                self.gather = ogd.PeerGather(env.ctx, self.shape, dst=0, slots=2,
                                             dst_only=not os.environ.get('OGN_BENCH_GATHER_EARLY'))
                if not os.environ.get('OGN_BENCH_GATHER_EARLY'):
                    t_ = self.tile
                    us = os.environ.get('OGN_BENCH_STAGGER_US')
                    self.gather.stagger(nz * (t_.y1 - t_.y0) * (t_.x1 - t_.x0) * 4,
                                        microseconds=float(us) if us is not None else None)
                self.gather_mode = ('owned correl tiles stored into rank 0 over NVLink peer memory '
                                    '(ogn_scatter_tile, CUDA IPC)')
            except Exception as exc:  # noqa: BLE001
                self.gather_mode = 'NCCL send/recv (peer mapping unavailable: %s)' % str(exc)[:120]
        self.thr_dev = torch.from_numpy(THRESHOLDS).to(dev)
        self.counts_dev = torch.zeros(2 * len(THRESHOLDS), dtype=torch.int64, device=dev)
        self.i = 0
        self.ar_events = []
        self.trace = {}
        self.last = {}
        self.lo = lib_origin
        self.ogd = ogd

    def _tick(self, name, t0):
        if self.env.args.trace:
            self.trace[name] = self.trace.get(name, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    def step(self):
        env, lo = self.env, self.lo
        i = self.i
        self.i = i + 1
        ny, nx = self.shape[1], self.shape[2]
        t0 = time.perf_counter()
        no_gather = os.environ.get('OGN_BENCH_NO_GATHER')
        if self.gather is not None and not no_gather:
            self.gather.attach(slot=i % 2)      # rank 0's own tile is stored by K2 itself
        # sync=False: nothing of the step comes back to the host (the extremum list lengths stay on the device),
        # so the host runs ahead and the GPU never idles between steps
        res = lo.step05(self.cube, self.fsf, None, self.profs, self.mask, 3, 1e-8, True,
                        out=self.out_sets[i % len(self.out_sets)], ctx=env.ctx,
                        tile=(self.tile, (ny, nx)) if self.world > 1 else None, sync=not ASYNC_STEP)
        t0 = self._tick('step05', t0)
        ext = res['extrema']                                                      # owned voxels, global indices
        n1, n0 = lo.purity_counts(ext, None, self.thr_dev, env.ctx, out=self.counts_dev)   # step06 counting loop
        t0 = self._tick('purity_counts', t0)
        if self.world > 1:
            if os.environ.get('OGN_TIMING_OFFSETS'):
                ea, eb = self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)
                ea.record()
                self.reducer.sum_(self.counts_dev)
                eb.record()
                self.ar_events.append((ea, eb))
            else:
                self.reducer.sum_(self.counts_dev)                                # NCCL allreduce of the histograms
            t0 = self._tick('allreduce', t0)
            # correl -> rank 0, enqueued last: the bulk stores would otherwise sit in front of the small
            # allreduce on the NVLink queues; this way they overlap the kernels of the next step instead
            if no_gather or (os.environ.get('OGN_BENCH_SKIP_LOCAL') and self.rank == 0):
                pass
            elif self.gather is not None:
                self.gather.scatter(res['correl'], self.tile, (ny, nx), slot=i % 2)
            else:
                self.last['correl_full'] = self.ogd.gather_owned(res['correl'], self.tile, self.all_tiles, self.shape)
            t0 = self._tick('gather', t0)
        self.last.update(res=res, ext=ext, slot=i % 2)
        return res

    def sync_all(self):
        import torch.distributed as dist
        self.torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
            self.torch.cuda.synchronize()

    def timed(self, steps, warmup, sample_clocks=False):
        """W warm-up steps, then K steps between CUDA events; max over ranks.  Returns a dict."""
        import torch.distributed as dist
        torch, env = self.torch, self.env
        for _ in range(warmup):
            self.step()
        self.sync_all()
        env.ctx.timing(True)
        env.ctx.timing_report()
        self.trace.clear()
        launches0 = env.ctx.launch_count
        sampler = None
        if sample_clocks:
            sampler = ClockSampler(env.local_rank)
            sampler.start()
            time.sleep(0.25)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.sync_all()
        self.ar_events = []
        e0.record()
        for _ in range(steps):
            self.step()
        if self.gather is not None:
            self.gather.join()                # the timed region ends after the last peer copy
        e1.record()
        self.sync_all()
        ms_total = e0.elapsed_time(e1)
        if env.args.trace:
            sys.stderr.write('rank %d host ms/step: %s\n' % (self.rank, ' '.join(
                '%s=%.3f' % (k, v * 1e3 / steps) for k, v in self.trace.items())))
        clocks = sampler.stop() if sampler else None
        launches = env.ctx.launch_count - launches0
        folded = env.ctx.fsf_folded
        stages = {}
        timeline = []
        for name, ms in env.ctx.timing_report():
            if '@' in name:                      # OGN_TIMING_OFFSETS=1: "stage@start_ms" -> per-rank timeline on stderr
                name, off = name.split('@')
                timeline.append((float(off), name, ms))
            stages.setdefault(name, []).append(ms)
        if timeline:             # development aid: the raw per-rank timeline (stage, start, duration) as JSON
            ar = [(e0.elapsed_time(a), a.elapsed_time(b)) for a, b in self.ar_events]
            os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
            with open(os.path.join(ROOT, 'gpurun_out', 'timeline_rank%d.json' % self.rank), 'w') as f:
                json.dump(dict(stages=sorted(timeline), allreduce=ar), f)
        env.ctx.timing(False)
        per_rank = None
        if self.world > 1:
            t = torch.tensor([ms_total], device=env.dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_total = float(t.item())
            lt = torch.tensor([launches], device=env.dev, dtype=torch.int64)
            dist.all_reduce(lt, op=dist.ReduceOp.SUM)
            launches = int(lt.item())
            per_rank = [None] * self.world
            dist.all_gather_object(per_rank, {k: round(float(np.mean(v)), 4) for k, v in stages.items()})
        nz, ny, nx = self.shape
        ms_step = ms_total / steps
        units = nz * ny * nx * self.nprof / 1e9
        return dict(ms_per_step=ms_step, value=units / (ms_step * 1e-3), units=units, launches=launches,
                    stages=stages, folded=folded, clocks=clocks, per_rank_stages=per_rank)

    # ---- end to end through the host API with pinned host buffers ------------------------------
    def e2e(self, steps):
        """step05 -> step06 counts -> step07 extraction through the host API, as the fused step mirror calls it
        (origin_b200.steps._run_compute_tglr) in a session that packed its mask once (origin_b200.steps.pack_mask) and
        keeps its buffers page-locked: float32 cube and bit-packed mask in pinned HOST memory in; correl,
        the two maps, the extremum lists, the per-threshold counts and the detection rows back on the HOST;
        correl_min and profile stay on the GPU behind lazy step products (fetched only if somebody reads them)."""
        import torch.distributed as dist
        from origin_b200 import _lib
        torch, env, lo = self.torch, self.env, self.lo
        shape = tuple(self.cube.shape)
        ty, tx = shape[1], shape[2]
        ny, nx = self.shape[1], self.shape[2]
        cube_h = _lib.pinned_empty(shape, np.float32)
        cube_h[...] = self.cube.cpu().numpy()
        cap = self.cap
        single = self.world == 1
        mask_np = self.mask.cpu().numpy()
        if single:           # packed once per session (np.packbits), uploaded every step
            packed = np.packbits(mask_np.reshape(-1))
            mask_h = _lib.pinned_empty(packed.shape, np.uint8)
            mask_h[...] = packed
        else:                # tile mode takes the byte mask
            mask_h = _lib.pinned_empty(shape, np.uint8)
            mask_h[...] = mask_np
        out_h = dict(correl=_lib.pinned_empty(shape, np.float32), maxmap=_lib.pinned_empty((ty, tx), np.float32),
                     minmap=_lib.pinned_empty((ty, tx), np.float32),
                     max_index=_lib.pinned_empty((cap,), np.int64), max_value=_lib.pinned_empty((cap,), np.float32),
                     min_index=_lib.pinned_empty((cap,), np.int64), min_value=_lib.pinned_empty((cap,), np.float32))
        dev = torch.device('cuda', env.ctx.device)
        # device-resident products are allocated once, like the step mirror's (they are outputs, not traffic)
        out_h['correl_min'] = torch.empty(shape, dtype=torch.float32, device=dev)
        out_h['profile'] = torch.empty(shape, dtype=torch.uint8, device=dev)
        tile_arg = (self.tile, (ny, nx)) if not single else None

        def step_e2e():
            if single:
                res = lo.step05(cube_h, self.fsf_host, None, self.profs, None, 3, 1e-8, True, out=out_h, ctx=env.ctx,
                                mask_bits=mask_h)
            else:
                res = lo.step05(cube_h, self.fsf_host, None, self.profs, mask_h, 3, 1e-8, True, out=out_h, ctx=env.ctx,
                                tile=tile_arg)
            n1, n0 = lo.purity_counts(res['extrema'], None, THRESHOLDS, env.ctx)            # step06 counting loop
            if not single:
                self.reducer.sum(np.concatenate([n1, n0]))
            # step07 extraction (tile mode: the lists carry whole-field indices, the profile lookup belongs to rank 0)
            rows = lo.threshold_rows(res['extrema'], 8.0, res['profile'] if single else None, 'max', env.ctx)
            return res, rows

        step_e2e()
        self.sync_all()
        n_e2e = max(2, min(steps, 5))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n_e2e):
            res, rows = step_e2e()
        e1.record()
        self.sync_all()
        wall = (time.perf_counter() - t0) * 1e3 / n_e2e
        ms_e2e = max(e0.elapsed_time(e1) / n_e2e, wall)
        if self.world > 1:
            t = torch.tensor([ms_e2e], device=env.dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e2e = float(t.item())
        n1c, n0c = res['extrema'].counts
        h2d = cube_h.nbytes + mask_h.nbytes + self.fsf_host.nbytes
        d2h = (out_h['correl'].nbytes + out_h['maxmap'].nbytes + out_h['minmap'].nbytes + 12 * (n1c + n0c) + 16
               + 13 * len(rows['z0']) + 8 * 2 * len(THRESHOLDS))
        nz, gny, gnx = self.shape
        units = nz * gny * gnx * self.nprof / 1e9
        return dict(value=units / (ms_e2e * 1e-3), unit='Gvoxel.profiles/s', ms_per_step=ms_e2e,
                    h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h), steps=n_e2e,
                    api='origin_b200.lib_origin.step05 -> %s (pinned host cube + %s mask in; correl, maps, lists out) + '
                        'purity_counts%s + threshold_rows'
                        % ('ogn_step05_bits' if single else 'ogn_step05_tile', 'bit-packed' if single else 'byte',
                           '' if single else ' + allreduce'),
                    lazy_on_device=['correl_min', 'profile', 'cube_local_max', 'cube_local_min'],
                    lazy_note='these step products are fetched from the GPU only when read (origin_b200.steps.LazyProduct); '
                              'fetching correl_min and profile as well adds %.2f GB of D2H per step'
                              % ((self.cube.numel() * 5) / 1e9),
                    n_detections=int(len(rows['z0'])), bytes_note='per rank' if self.world > 1 else None)

    # ---- parity on the benchmark cube itself -----------------------------------------------------
    def spot_points(self, ext_host, n_random=24, n_list=16):
        nz, ny, nx = self.shape
        rng = np.random.default_rng(2024)
        pts = []
        for z in (0, nz - 1):                                   # the 8 corners
            for y in (0, ny - 1):
                for x in (0, nx - 1):
                    pts.append((z, y, x))
        mid = (nz // 2, ny // 2, nx // 2)
        for ax in range(3):                                      # edge / face centres
            for lo_hi in (0, 1):
                p = list(mid)
                p[ax] = 0 if lo_hi == 0 else self.shape[ax] - 1
                pts.append(tuple(p))
                q = list(p)
                q[(ax + 1) % 3] = 0
                pts.append(tuple(q))
        for t in self.all_tiles:                                 # both sides of every tile seam
            for (y, x) in ((t.y0, (t.x0 + t.x1) // 2), (t.y0 - 1, (t.x0 + t.x1) // 2),
                           ((t.y0 + t.y1) // 2, t.x0), ((t.y0 + t.y1) // 2, t.x0 - 1), (t.y0, t.x0), (t.y0 - 1, t.x0 - 1)):
                if 0 <= y < ny and 0 <= x < nx and len(self.all_tiles) > 1:
                    pts.append((int(rng.integers(0, nz)), int(y), int(x)))
        for _ in range(n_random):
            pts.append((int(rng.integers(0, nz)), int(rng.integers(0, ny)), int(rng.integers(0, nx))))
        for key, n in (('max', n_list), ('min', n_list // 2)):   # voxels the GPU flagged, strongest first
            idx, val = ext_host[key]
            if len(idx):
                order = np.argsort(-val)[:n // 2]
                pick = np.concatenate([order, rng.integers(0, len(idx), n - len(order))])
                for i in pick:
                    pts.append(tuple(int(v) for v in np.unravel_index(int(idx[i]), self.shape)))
        seen, out = set(), []
        for p in pts:
            if p not in seen:
                seen.add(p)
                out.append(p)
        return out

    def parity_spot(self, correl, correl_min, profile, ext_host, full_cube):
        """Float64 direct-space oracle on seeded voxels of the benchmark cube and their 3x3x3 neighbourhoods.
        ``correl`` / ``correl_min`` / ``profile`` / ``full_cube`` are device cubes of the whole field;
        ``ext_host`` = {'max': (index, value), 'min': (index, value)} global lists."""
        from oracle import origin_oracle as orc
        from origin_b200 import synthetic
        torch = self.torch
        nz, ny, nx = self.shape
        pts = self.spot_points(ext_host)
        rms = float(torch.sqrt(torch.mean(correl.double() ** 2)).item())
        half = max((len(d) - 1) // 2 for d in self.prof_cut)
        c = PSF_SIZE // 2
        worst = 0.0
        stats = dict(n=len(pts), argmax_ties=0, argmax_mismatches=0, extremum_ties=0, extremum_mismatches=0,
                     maxima_checked=0, minima_checked=0, input_mismatch_voxels=0)
        max_set, min_set = ext_host['max'][0], ext_host['min'][0]
        for (z, y, x) in pts:
            z0, z1 = max(0, z - 1 - half - 1), min(nz, z + 2 + half + 1)
            y0, y1 = max(0, y - 1 - c), min(ny, y + 2 + c)
            x0, x1 = max(0, x - 1 - c), min(nx, x + 2 + c)
            win, mwin = synthetic.bench_window(self.shape, ((z0, z1), (y0, y1), (x0, x1)), self.fsf_host, seed=0)
            # the oracle runs on the DEVICE cube's values; the host generator must agree with it (float64 libm vs
            # CUDA log / cos: about one voxel in 1e9 may differ in the last place)
            dev_win = full_cube[z0:z1, y0:y1, x0:x1].cpu().numpy()
            stats['input_mismatch_voxels'] += int(np.count_nonzero(dev_win != win))
            win = dev_win
            box = orc.spot_box(win, self.fsf_host, self.prof_cut, (z0, y0, x0), self.shape, (z, y, x))
            tk = box['tk']                                     # [3][3][3][K], NaN outside the cube
            valid = box['valid']
            mloc = np.zeros((3, 3, 3), dtype=bool)
            for dz in range(3):
                for dy in range(3):
                    for dx in range(3):
                        zz, yy, xx = z - 1 + dz, y - 1 + dy, x - 1 + dx
                        if valid[dz, dy, dx]:
                            mloc[dz, dy, dx] = bool(mwin[zz - z0, yy - y0, xx - x0])
            cref = np.where(valid, np.nanmax(np.where(valid[..., None], tk, -np.inf), axis=-1), -np.inf)
            mref = np.where(valid, np.nanmin(np.where(valid[..., None], tk, np.inf), axis=-1), np.inf)
            cref_masked = np.where(mloc, 0.0, cref)            # steps.py:781: correl zeroed under the mask
            t_c = tk[1, 1, 1]
            g_c = float(correl[z, y, x].item())
            g_m = float(correl_min[z, y, x].item())
            g_p = int(profile[z, y, x].item())
            for got, ref in ((g_c, cref_masked[1, 1, 1]), (g_m, mref[1, 1, 1])):
                bound = RTOL * max(abs(ref), rms)
                worst = max(worst, abs(got - ref) / bound)
            tol = RTOL * max(abs(cref[1, 1, 1]), rms)
            if not mloc[1, 1, 1]:
                k_ref = int(np.argmax(t_c))
                if g_p != k_ref:
                    if t_c[k_ref] - t_c[g_p] <= 2 * tol:
                        stats['argmax_ties'] += 1
                    else:
                        stats['argmax_mismatches'] += 1
            # local extremum: value == max over the clamped 3x3x3 window and not masked (lib_origin.py:1244-1247)
            lin = (z * ny + y) * nx + x
            for arr, lst, key in ((cref_masked, max_set, 'maxima_checked'), (-mref, min_set, 'minima_checked')):
                centre = arr[1, 1, 1]
                others = np.where(valid, arr, -np.inf).copy()
                others[1, 1, 1] = -np.inf
                nb = others.max()
                ref_is = (centre >= nb) and not mloc[1, 1, 1]
                j = np.searchsorted(lst, lin)
                got_is = bool(j < len(lst) and lst[j] == lin)
                stats[key] += int(got_is)
                if ref_is != got_is:
                    if abs(centre - nb) <= 2 * tol and not mloc[1, 1, 1]:
                        stats['extremum_ties'] += 1
                    else:
                        stats['extremum_mismatches'] += 1
        stats['worst_over_bound'] = worst
        stats['rms_correl'] = rms
        stats['bound'] = '|d| <= %g * max(|ref|, rms(correl))' % RTOL
        stats['oracle'] = ('oracle.origin_oracle.spot_box: float64 direct-space T_k on the 3x3x3 neighbourhood of each '
                           'voxel (SURVEY appendix A.2), inputs cut from the benchmark cube')
        return stats

    def host_lists(self, ext):
        return dict(max=(ext.max_index.cpu().numpy(), ext.max_value.cpu().numpy()),
                    min=(ext.min_index.cpu().numpy(), ext.min_value.cpu().numpy()))

    def sharded_parity(self):
        """N > 1: rank 0 recomputes the whole cube on its own GPU and compares it with the peer-gathered cube of
        the last step, the merged per-rank extremum lists and the all-reduced purity counts."""
        import torch.distributed as dist
        from origin_b200 import synthetic
        torch, env, lo = self.torch, self.env, self.lo
        nz, ny, nx = self.shape
        ext = self.last['ext']
        if self.gather is not None:
            self.gather.wait()
        # per-rank lists -> rank 0 (padded to the longest)
        n = torch.tensor(list(ext.counts), device=env.dev, dtype=torch.int64)
        sizes = [torch.zeros_like(n) for _ in range(self.world)]
        dist.all_gather(sizes, n)
        cap = int(max(int(s.max().item()) for s in sizes))

        def padded(t, fill):
            buf = torch.full((cap,), fill, dtype=t.dtype, device=env.dev)
            buf[:len(t)] = t
            return buf
        gathered = {}
        for key, fill in (('max_index', -1), ('max_value', 0.0), ('min_index', -1), ('min_value', 0.0)):
            bufs = [torch.empty(cap, dtype=getattr(ext, key).dtype, device=env.dev) for _ in range(self.world)]
            dist.all_gather(bufs, padded(getattr(ext, key), fill))
            gathered[key] = bufs
        counts_all = self.counts_dev.clone()
        out = None
        if self.rank == 0:
            cube_g, mask_g = synthetic.bench_window(self.shape, None, self.fsf_host, seed=0, xp=torch, device=env.dev)
            full = lo.step05(cube_g, self.fsf, None, self.profs, mask_g, 3, 1e-8, True, ctx=env.ctx)
            fext = full['extrema']
            if self.gather is not None:
                got = self.gather.result(self.last['slot'])
            else:
                got = self.last['correl_full']
            dmax = float((got - full['correl']).abs().max().item())
            same = {}
            for which, kn in (('max', 0), ('min', 1)):
                idx = torch.cat([b[:int(s[kn].item())] for b, s in zip(gathered[which + '_index'], sizes)])
                val = torch.cat([b[:int(s[kn].item())] for b, s in zip(gathered[which + '_value'], sizes)])
                order = torch.argsort(idx)
                ref_i, ref_v = getattr(fext, which + '_index'), getattr(fext, which + '_value')
                same[which] = bool(len(idx) == len(ref_i) and torch.equal(idx[order], ref_i)
                                   and torch.equal(val[order], ref_v))
            n1, n0 = lo.purity_counts(fext, None, THRESHOLDS, env.ctx)
            ref_counts = torch.from_numpy(np.concatenate([n1, n0])).to(env.dev)
            out = dict(correl_max_abs=dmax, lists_identical=bool(same['max'] and same['min']),
                       counts_identical=bool(torch.equal(ref_counts, counts_all)),
                       n_local_extrema=[int(v) for v in fext.counts],
                       what='rank 0 recomputed the whole %dx%dx%d cube on one GPU after the timed region; compared with '
                            'the %s cube of the last step, the merged per-rank extremum lists (index and value) and the '
                            'all-reduced per-threshold counts' % (nz, ny, nx, 'peer-gathered' if self.gather else 'NCCL-gathered'))
            self.last['full'] = full
            self.last['full_cube'] = cube_g
        dist.barrier()
        return out

    def close(self):
        if self.gather is not None:
            self.sync_all()
            self.gather.close()
            self.gather = None


class Env:
    pass


def run_c3(env, hbm_peak, fp32_peak, steps, warmup):
    """north_star config 3: the same cube with Dico_FWHM_2_12 plus step01 (DCT continuum subtraction +
    standardisation + local extrema of cube_std), all device-resident, CUDA events."""
    import torch
    from origin_b200 import lib_origin as lo, synthetic
    job = Job(env, env.args.shape, '2_12', world=1)
    r = job.timed(steps, warmup)
    nz, ny, nx = job.shape
    vol = nz * ny * nx
    models = kernel_models(job.prof_cut, r['folded'], job.nprof > 3)
    roofs, dominant = stage_rooflines(r['stages'], models, vol, fp32_peak, hbm_peak, {})
    # step01 input: the noise cube scaled by a smooth variance cube, plus a smooth continuum (raw 0 / var inf
    # under the mask, origin.py:262-274)
    lam = torch.linspace(0.0, 1.0, nz, device=env.dev)
    var = (1.0 + 0.3 * torch.sin(6.0 * lam) ** 2)[:, None, None] * (1.0 + 0.25 * torch.cos(
        torch.linspace(0, 6.28, ny, device=env.dev))[:, None] * torch.cos(torch.linspace(0, 6.28, nx, device=env.dev))[None, :])[None]
    var = var.to(torch.float32).contiguous()
    raw = job.cube * torch.sqrt(var) + (50.0 * (0.6 + 0.4 * lam))[:, None, None]
    m = job.mask.bool()
    raw[m] = 0.0
    var[m] = float('inf')
    del m
    ctx = env.ctx

    def step01():
        out = lo.preprocess(raw, var, job.mask, 10, False, ctx=ctx)
        ext, _, _ = lo.local_extrema(out['cube_std'], out['cube_std'], job.mask, 3, capacity=max(4096, vol // 16), ctx=ctx)
        return out, ext

    for _ in range(3):          # same binding pattern as the timed loop: two product sets end up in the allocator's cache
        out, ext = step01()
    torch.cuda.synchronize()
    ctx.timing(True)
    ctx.timing_report()
    n01 = max(2, min(steps, 5))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n01):
        out, ext = step01()
    e1.record()
    torch.cuda.synchronize()
    ms01 = e0.elapsed_time(e1) / n01
    st = {}
    for name, ms in ctx.timing_report():
        st.setdefault(name, []).append(ms)
    ctx.timing(False)
    st = {k: float(np.mean(v)) for k, v in st.items()}
    k5 = sum(v for k, v in st.items() if k.startswith('k5'))          # DCT fit, per-lambda sums, standardisation
    step01_roof = dict(
        kernels_ms=st, dct_kernels_ms=k5, bound='hbm', bytes_per_voxel=17.0,
        achieved=17.0 * vol / (k5 * 1e-3) / 1e9 if k5 else None, peak=hbm_peak, unit='GB/s',
        frac=(17.0 * vol / (k5 * 1e-3) / 1e9 / hbm_peak) if (k5 and hbm_peak) else None,
        fp64=dict(dfma_per_voxel=65, tflops=2 * 65.0 * vol / (k5 * 1e-3) / 1e12 if k5 else None),
        note='raw 4 + var 4 + mask 1 in, cube_std 4 + cont_dct 4 out = 17 B/voxel algorithmic (SURVEY 8d)')
    res = dict(
        workload='step01 (DCT order 10, weighted) + step05 TGLR + local extrema + step06 counts, %dx%dx%d, '
                 'Dico_FWHM_2_12 (%d profiles)' % (nz, ny, nx, job.nprof),
        ms_per_step=r['ms_per_step'], value=r['value'], unit='Gvoxel.profiles/s', steps=steps, warmup=warmup,
        step01_ms=ms01, step01_wall_includes='ogn_preprocess (fit, per-lambda mean on the device, standardisation, image reductions) '
                                             '+ local extrema of cube_std',
        total_ms=r['ms_per_step'] + ms01, gpu_launches=r['launches'],
        roofline=dict(roofs[dominant], dominant_stage=dominant) if dominant else None,
        kernels=roofs, step01_roofline=step01_roof,
        other_ms={k: float(np.mean(v)) for k, v in r['stages'].items() if k not in roofs and k != 'step05_span'},
        n_local_extrema_std=[int(v) for v in ext.counts])
    job.close()
    del raw, var, out
    torch.cuda.empty_cache()
    return res


def run_next_rows(env):
    """SURVEY 8(f) rows 1 and 4 in the driver's run: the step04 greedy PCA of one 3681x96x96 area with twelve continuum
    sources left in it, and the step08 line estimation of 32 detections (288 windows of 3681x25x25) in it - the
    workload of tools/pca_lines_probe.py - device-resident, wall clock between synchronisations (both loops are
    host-driven: the Lanczos restarts read the tridiagonal matrices back).  The unmodified reference on the same
    inputs on one host core is in profiles/r02_pca_lines_probe_cpu_reference.json (13.3 s; 3.5 s per detection);
    parity is what tests/test_gpu_pca.py and tests/test_gpu_lines.py check against the reference's own functions."""
    import torch
    from origin_b200 import lib_origin as lo, synthetic
    nz, ny, nx = SHAPE[0], 96, 96
    rng = np.random.default_rng(11)
    fsf = synthetic.moffat_fsf(nz)
    cube = rng.standard_normal((nz, ny, nx)).astype(np.float32)
    lam = np.linspace(0, 1, nz)
    for _ in range(12):
        y0, x0 = int(rng.integers(12, ny - 12)), int(rng.integers(12, nx - 12))
        spec = rng.uniform(2, 8) * (0.6 + 0.4 * np.cos(rng.uniform(1, 6) * lam + rng.uniform(0, 3)))
        cube[:, y0 - 12:y0 + 13, x0 - 12:x0 + 13] += (spec[:, None, None] * fsf / fsf.max(axis=(1, 2), keepdims=True)).astype(np.float32)
    areamap = np.ones((ny, nx), dtype=int)
    dets = dict(z0=rng.integers(50, nz - 50, 32), y0=rng.integers(0, ny, 32), x0=rng.integers(0, nx, 32))
    var = (1.0 + 0.3 * np.sin(6 * lam) ** 2)[:, None, None] * np.ones((1, ny, nx))
    raw = torch.from_numpy((cube * np.sqrt(var)).astype(np.float32)).to(env.dev)
    var = torch.from_numpy(var.astype(np.float32)).to(env.dev)
    c = torch.from_numpy(cube).to(env.dev)
    ctx = env.ctx
    launches0 = ctx.launch_count
    test, _, _, thr, _, _ = lo.Compute_PCA_threshold(c.reshape(nz, -1), 0.01)
    out = {}
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        faint, map_o2, nstop = lo.Compute_GreedyPCA_area(1, c, areamap, 50, [thr], 100, [test], ctx=ctx)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    out['step04_greedy_pca'] = dict(
        workload='one %dx%dx%d area, 12 continuum sources' % (nz, ny, nx), seconds=dt, iterations=float(map_o2.max()),
        nuisance_spaxels=int((map_o2 > 0).sum()), nstop=int(nstop), cube_faint_on_device=bool(faint.is_cuda),
        reference_seconds_one_core=13.35, reference_from='profiles/r02_pca_lines_probe_cpu_reference.json')
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        cat2, lin_est, _ = lo.estimation_line(dets, raw, var, fsf, ctx=ctx)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    out['step08_line_estimation'] = dict(
        workload='32 detections, grid of 9 offsets each = 288 windows of %dx25x25, order_dct 30' % nz, seconds=dt,
        ms_per_detection=dt * 1e3 / 32, finite_lines=int(sum(np.isfinite(l).all() for l in lin_est)),
        reference_ms_per_detection_one_core=3489.5, reference_from='profiles/r02_pca_lines_probe_cpu_reference.json')
    out['gpu_launches'] = int(ctx.launch_count - launches0)
    del raw, var, c, faint
    torch.cuda.empty_cache()
    return out


def run_c5(env, steps, warmup):
    """north_star config 5: the 3681x900x900 mosaic-sized cube with Dico_FWHM_2_12 on all ranks, and on rank 0
    alone (1-GPU time), so that the speed-up is timed in one driver run."""
    import torch
    import torch.distributed as dist
    job = Job(env, C5_SHAPE, '2_12')
    r = job.timed(steps, warmup)
    res = dict(workload='step05 TGLR + local extrema + step06 counts + allreduce + correl gather, %dx%dx%d, '
                        'Dico_FWHM_2_12' % C5_SHAPE, n_gpus=env.world, ms_per_step=r['ms_per_step'], value=r['value'],
               unit='Gvoxel.profiles/s', steps=steps, warmup=warmup, gpu_launches=r['launches'])
    job.close()
    del job
    torch.cuda.empty_cache()
    one = None
    if env.rank == 0:
        job1 = Job(env, C5_SHAPE, '2_12', world=1)
        n1 = max(2, min(steps, 3))
        r1 = job1.timed(n1, 1)
        one = dict(ms_per_step=r1['ms_per_step'], value=r1['value'], steps=n1, warmup=1)
        job1.close()
        del job1
        torch.cuda.empty_cache()
    dist.barrier()
    if env.rank == 0:
        res['one_gpu'] = one
        res['speedup_vs_one_gpu'] = one['ms_per_step'] / res['ms_per_step']
    return res


def main_gpu(args):
    import torch
    import torch.distributed as dist
    from origin_b200 import _lib

    env = Env()
    env.args = args
    env.world = int(os.environ.get('WORLD_SIZE', '1'))
    env.rank = int(os.environ.get('RANK', '0'))
    env.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the ported path has no CPU fallback')
    torch.cuda.set_device(env.local_rank)
    if env.world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', env.local_rank))
    env.dev = torch.device('cuda', env.local_rank)
    env.ctx = _lib.default_context(env.local_rank)
    world, rank = env.world, env.rank

    job = Job(env, args.shape, args.dico)
    nz, ny, nx = job.shape
    r = job.timed(args.steps, args.warmup, sample_clocks=True)

    e2e = None
    if not args.no_e2e:
        e2e = job.e2e(args.steps)

    # ---- parity on the cube that was just timed -------------------------------------------------
    sharded = None
    if world > 1 and not args.no_parity:
        sharded = job.sharded_parity()
    spot = None
    if rank == 0 and not args.no_parity:
        if world == 1:
            res, full_cube = job.lo.step05(job.cube, job.fsf, None, job.profs, job.mask, 3, 1e-8, True, ctx=env.ctx), job.cube
        else:
            # the checked products are rank 0's single-GPU recompute, which sharded_parity just compared with the
            # gathered cube / merged lists of the multi-GPU step; the points include both sides of every tile seam
            res, full_cube = job.last['full'], job.last['full_cube']
        t0 = time.perf_counter()
        spot = job.parity_spot(res['correl'], res['correl_min'], res['profile'], job.host_lists(res['extrema']), full_cube)
        spot['seconds'] = time.perf_counter() - t0
        del res, full_cube
    job.last.clear()

    fp32 = fma_peak() if rank == 0 else {}
    fp32_peak = fp32.get('fp32_tflops')
    hbm_peak = _load_json('MEASURED_PEAKS.json').get('hbm_gbs')
    summ = _load_json('profiles', 'ncu_summary.json')
    job.close()
    cube_gb = job.cube.numel() * 4 / 1e9
    vol_tile, gather_mode, nprof, prof_cut = job.vol_tile, job.gather_mode, job.nprof, job.prof_cut
    del job
    torch.cuda.empty_cache()

    # ---- the other north_star configurations, timed in the same driver run ------------------------
    configs = {}
    if not args.no_configs and tuple(args.shape) == SHAPE and args.dico == '3FWHM':
        csteps, cwarm = max(3, min(args.steps, 10)), 3
        if world == 1:
            configs['c3'] = run_c3(env, hbm_peak, fp32_peak, csteps, cwarm)
            try:                        # the "next" rows of SURVEY 8(f); never at the expense of the line itself
                configs['next_rows'] = run_next_rows(env)
            except Exception as exc:  # noqa: BLE001
                configs['next_rows'] = dict(error=repr(exc)[:300])
        if world == 8:
            c5 = run_c5(env, max(3, min(args.steps, 5)), 2)
            if rank == 0:
                configs['c5'] = c5

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -----------------------------------------------------------
    models = kernel_models(prof_cut, r['folded'], nprof > 3)
    traffic = {k: summ.get(k + '_dram_bytes_per_launch') for k in models}   # ncu dram__bytes_read+write per launch
    if nprof > 3:
        traffic['k2_spectral_glr'] = summ.get('k2f_folded_glr_dram_bytes_per_launch')
    if tuple(args.shape) != SHAPE:
        traffic = {}                                     # the captures are of the 3681x320x320 cube
    roofs, dominant = stage_rooflines(r['stages'], models, vol_tile, fp32_peak, hbm_peak, traffic if world == 1 else {})
    roofline = None
    if dominant:
        roofline = dict(roofs[dominant], dominant_stage=dominant,
                        peak_source=('FP32 FFMA peak measured on this device in this run by tools/fma_peak '
                                     '(MEASURED_PEAKS.json has no FP32 figure; nominal 148 SM x 128 lanes x 2 x 1.965 GHz '
                                     '= 74.4)' if roofs[dominant]['bound'] == 'fp32' else
                                     'MEASURED_PEAKS.json hbm_gbs'),
                        frac_meaning='achieved = FP32 issue slots the kernel EXECUTES (FFMA, FADD, FMUL each count 2 flop '
                                     'slots) / time: the pipe utilisation, <= 1; achieved_algorithmic = the flops of the '
                                     "direct form (SURVEY.md 8d) / time, which can exceed the peak when the kernel folds "
                                     'symmetric taps' if roofs[dominant]['bound'] == 'fp32' else None,
                        note='the schema value "tensor" does not apply: step05 is FP32-FMA bound (SURVEY.md 8d); tensor '
                             'cores cannot meet the 1e-5 parity bound in one pass')
    total_flops = (2.0 * PSF_SIZE ** 2 + 2.0 * sum(len(p) for p in prof_cut)) * nz * ny * nx
    ms_step = r['ms_per_step']
    cfg = workload_config(args.shape, args.dico, nprof, world)
    line = dict(
        metric='step05 TGLR throughput', value=r['value'], unit='Gvoxel.profiles/s', n_gpus=world, steps=args.steps,
        warmup=args.warmup, ms_per_step=ms_step, higher_is_better=True, scaling='strong', vs_baseline=None,
        dtype='f32', data='synthetic', config=cfg, gather=gather_mode,
        timed_region='CUDA events on the launching stream around %d steps, barrier + synchronize on both sides, '
                     'max over ranks; %.1f GB of inputs per rank' % (args.steps, cube_gb),
        e2e=e2e, gpu_launches=int(r['launches']), clocks=r['clocks'], roofline=roofline, kernels=roofs,
        other_ms={k: float(np.mean(v)) for k, v in r['stages'].items() if k not in roofs and k != 'step05_span'},
        step05_span_ms=float(np.mean(r['stages'].get('step05_span', [np.nan]))),
        step_fp32_tflops=total_flops / (ms_step * 1e-3) / 1e12,
        step_fp32_frac=(total_flops / (ms_step * 1e-3) / 1e12 / fp32_peak) if fp32_peak else None,
        fma_peak=fp32, parity_spot=spot, sharded_parity=sharded, configs=configs or None,
        per_rank_stage_ms=r['per_rank_stages'],
    )
    if world == 1 and not args.no_cpu:
        # the reference arm in a child process (no fork of a CUDA process), one bounded step
        try:
            cp = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', '1',
                                 '--warmup', '0', '--dico', args.dico, '--shape'] + [str(v) for v in args.shape],
                                capture_output=True, text=True, timeout=600,
                                env={k: v for k, v in os.environ.items() if k != 'OMP_NUM_THREADS'})
            cb = json.loads(cp.stdout.strip().splitlines()[-1])['cpu_baseline']
            cb['seconds_for_sample'] = json.loads(cp.stdout.strip().splitlines()[-1])['ms_per_step'] / 1e3
        except Exception as exc:  # noqa: BLE001
            cb = dict(error=str(exc)[:200])
        line['cpu_baseline'] = cb
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    a = parse_args()
    if a.impl == 'reference':
        main_reference(a)
    else:
        main_gpu(a)
