#!/usr/bin/env python
"""
bench.py -- output voxels/s of the dense Lucas-Kanade hot path (calc_flow3D / calc_flow2D).

    python bench.py --gpus N --steps K --warmup W [--workload cfg4] [--precision fp64]
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one pass of the hot path over the workload's whole synthetic time-lapse: every output
timepoint (Nt - 2*ceil(3 tSig) of them) is computed once, sharded by output timepoint across the N
ranks (strong scaling; no data-path collective).  `value` = output voxels of all ranks / max-over-ranks
device time, inputs resident in HBM.  `e2e` = the same metric through the streaming engine under
process_flow (timelapse.FlowStream: host frames in, host results out, host<->device copies inside the
timed region, a bounded number of timepoints of the same volume shape); beside it
`e2e.calc_flow_call_value` / `calc_flow_call_plain_numpy_value`: the synchronous drop-in call
calc_flow3D(host window) -> host arrays with pinned / ordinary NumPy buffers.  `roofline` = the dominant
kernel, timed live by the library's per-stage event brackets (of3d_set_profile), with every stage in
`roofline.stages` and the whole pipeline in `roofline.pipeline`.  `cpu_baseline` / `--impl reference` = the reference's
NumPy/SciPy algorithm (oracle port using the reference's own scipy.ndimage.correlate1d and
numpy.linalg.eigvals calls) on the box's host cores, on a z/y/x-cropped sample of the same workload.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json configs; uint16 camera-like input
WORKLOADS = {
    'cfg1': dict(shape=(7, 32, 128, 128), sig=(1, 1, 4)),
    'cfg2': dict(shape=(31, 2048, 2048), sig=(1.5, 1, 4)),
    'cfg3': dict(shape=(31, 64, 512, 512), sig=(3, 2, 6)),
    'cfg4': dict(shape=(61, 128, 1024, 1024), sig=(3, 1, 4)),   # the configuration the metric is quoted on
    # one 19-frame window of a 2048x2048x512 volume: too large for one GPU, sharded by z-slab with an NCCL halo
    # exchange (needs --gpus >= 4 in fp64, >= 2 in fp32)
    'cfg5': dict(shape=(19, 512, 2048, 2048), sig=(3, 3, 8), zslab=True),
    # (debug) the kernels of cfg5 on one GPU: a 112-plane slab -- 64 owned planes + 2 x 24 window planes -- of the cfg5 volume
    'cfg5slab': dict(shape=(21, 112, 2048, 2048), sig=(3, 3, 8)),
}
METRIC = 'output voxels/s (vx,vy,vz,rel) 1024x1024x128 stack'
UNIT = 'voxels/s'
METRICS = {
    'cfg1': 'output voxels/s (vx,vy,vz,rel) 128x128x32 volume, 7 timepoints',
    'cfg2': 'output pixels/s (vx,vy,rel) 2048x2048 frame time-lapse (calc_flow2D)',
    'cfg3': 'output voxels/s (vx,vy,vz,rel) 512x512x64 stack, 31 timepoints',
    'cfg4': METRIC,
    'cfg5': 'output voxels/s (vx,vy,vz,rel) 2048x2048x512 volume, z-slab sharded',
    'cfg5slab': 'output voxels/s (vx,vy,vz,rel) 2048x2048x112 slab of the cfg5 volume (kernel study)',
}
REF_DIR = os.path.join(ROOT, 'baseline', '_ref')


def load_reference():
    """The UNMODIFIED reference module (baseline/_ref/calc_flow.py, copied verbatim from the reference checkout by
    __graft_entry__.build(); git-ignored, travels to the GPU box).  Its two I/O-only imports that are not installed here
    (tifffile, natsort) are stubbed; calc_flow2D / calc_flow3D do not use them.  None when the copy is absent."""
    path = os.path.join(REF_DIR, 'calc_flow.py')
    if not os.path.exists(path):
        return None
    import importlib.util
    import types
    for name, attrs in (('tifffile', {}), ('natsort', {'natsorted': sorted})):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                m.__dict__.update(attrs)
                sys.modules[name] = m
    spec = importlib.util.spec_from_file_location('of3d_reference_calc_flow', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def fma_peak(precision):
    """Measured CUDA-core FMA issue peak of this pool's B200 (tools/fma_peak.cu -> profiles/r01_fma_peak.json), TFMA/s."""
    best = {'fp64': 17.0, 'fp32': 36.2}          # values recorded in DESIGN.md; refreshed from the profile if present
    try:
        for line in open(os.path.join(ROOT, 'profiles', 'r01_fma_peak.json')):
            d = json.loads(line)
            if d.get('pipe') in best:
                best[d['pipe']] = max(best[d['pipe']], float(d['tfma_per_s']))
    except Exception:
        pass
    return best[precision]


def measured_traffic(workload, precision):
    """DRAM bytes per output voxel of the whole per-timepoint pipeline, from the committed ncu launch lists
    (profiles/r02_traffic.json: sum of dram__bytes_read.sum + dram__bytes_write.sum over the launches of one timepoint)."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'r02_traffic.json')) as fh:
            return json.load(fh).get('%s_%s' % (workload, precision))
    except Exception:
        return None


def alg_bytes_per_voxel(sig, ndim, precision, in_itemsize=2):
    """one calc_flow call reads its Kt-frame window once and writes its outputs once."""
    kt = 2 * math.ceil(3 * sig[1]) + 1
    return kt * in_itemsize + (ndim + 1) * (8 if precision == 'fp64' else 4)


def alg_fma_per_voxel(sig, ndim):
    """SURVEY.md 8(d) second roof: FMAs of the direct separable evaluation."""
    kr = 2 * math.ceil(3 * sig[0]) + 1
    ks = 2 * math.ceil(3 * sig[0] / 4) + 1
    kt = 2 * math.ceil(3 * sig[1]) + 1
    kw = 2 * math.ceil(3 * sig[2]) + 1
    if ndim == 3:
        return (kt - 1) + 3 * kr + 3 * kr + 6 * ks + 9 + 27 * kw + 120    # (the z-first order needs 13 fewer)
    return (kt - 1) + 2 * kr + 2 * kr + 2 * ks + 5 + 10 * kw + 30


def stage_model(sig, ndim, precision, in_itemsize=2, fused=True):
    """Per-stage algorithmic bytes and FMAs per voxel of the marching pipeline: every stage reads its
    inputs once and writes its outputs once; FMA counts are those of the direct separable evaluation the stage performs.
    3D runs z first: temporal derivative + three z filters in one kernel, then four in-plane kernels."""
    e = 8 if precision == 'fp64' else 4
    kr = 2 * math.ceil(3 * sig[0]) + 1
    ks = 2 * math.ceil(3 * sig[0] / 4) + 1
    kt = 2 * math.ceil(3 * sig[1]) + 1
    kw = 2 * math.ceil(3 * sig[2]) + 1
    if ndim == 3:
        m = {'gradient_z': ((kt * in_itemsize if fused else 2 * e) + 3 * e, (kt - 1 if fused else 0) + 2 * kr + ks),
             'gradient_xy': (4 * e + 4 * e, 4 * kr + 4 * ks),
             'products_window_z': (4 * e + 9 * e, 9 + 9 * kw),
             'window_xy_solve': (9 * e + 4 * e, 18 * kw + 120)}
        if not fused:
            m['temporal'] = (kt * in_itemsize + 2 * e, kt - 1)
        return m
    return {'temporal': (kt * in_itemsize + 2 * e, kt - 1),
            'gradient_xy': (3 * e + 3 * e, 4 * kr + 2 * ks),
            'window_xy_solve': (3 * e + 3 * e, 5 + 10 * kw + 30)}


def stage_report(stages, model, vol, timepoints, hbm_peak, fma_pk):
    """stages = Context.stage_times() of `timepoints` operator calls -> per-stage roofline rows, longest first"""
    rows = []
    for name, (ms, n) in stages.items():
        per_tp = ms / max(timepoints, 1)
        row = {'stage': name, 'ms_per_timepoint': per_tp, 'launches_per_timepoint': n / max(timepoints, 1)}
        if name in model and per_tp > 0:
            b, f = model[name]
            row.update({'alg_bytes_per_voxel': b, 'alg_gbs': b * vol / per_tp / 1e6, 'hbm_frac': b * vol / per_tp / 1e6 / hbm_peak,
                        'alg_fma_per_voxel': f, 'tfma_per_s': f * vol / per_tp / 1e9,
                        'fp_frac': (f * vol / per_tp / 1e9 / fma_pk) if fma_pk else None})
        rows.append(row)
    rows.sort(key=lambda r: -r['ms_per_timepoint'])
    return rows


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(',')]
                if len(f) < 8:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[4:8]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------ CPU arm
_REF = None


def cpu_kind():
    return 'reference' if os.path.exists(os.path.join(REF_DIR, 'calc_flow.py')) else 'port'


def _cpu_worker(args):
    """One calc_flow call of the reference itself (baseline/_ref) or, when that copy is absent, of the oracle port"""
    global _REF
    img, sig, ndim = args
    if _REF is None:
        _REF = load_reference() or False
    t0 = time.perf_counter()
    if _REF:
        (_REF.calc_flow3D if ndim == 3 else _REF.calc_flow2D)(img, *sig)
    else:
        from oracle import lk_oracle as orc
        if ndim == 3:
            orc.lk_flow3d(img, *sig, rel_mode='reference', use_scipy=True)
        else:
            orc.lk_flow2d(img, *sig, use_scipy=True)
    return time.perf_counter() - t0


def cpu_sample(workload, target_voxels, seed=0):
    """A cropped window of the workload (same sigmas, same dtype, same blob model) for the CPU arm."""
    from opticalflow3d_dev_b200.synth import make_stack
    w = WORKLOADS[workload]
    sig = w['sig']
    kt = 2 * math.ceil(3 * sig[1]) + 1
    sp = list(w['shape'][1:])
    while np.prod(sp) > target_voxels:      # halve the largest axis until it fits
        i = int(np.argmax(sp))
        sp[i] = max(8, sp[i] // 2)
        if all(s <= 8 for s in sp):
            break
    return make_stack((kt,) + tuple(sp), seed=1000 + seed, dtype=np.uint16), sig, len(sp)


def run_cpu(workload, cores, steps, warmup, target_voxels):
    """Reference algorithm on `cores` processes, one cropped window each per step."""
    import multiprocessing as mp
    jobs = [cpu_sample(workload, target_voxels, seed=i) for i in range(cores)]
    vox = sum(int(np.prod(j[0].shape[1:])) for j in jobs)
    times = []
    if cores == 1:
        for it in range(warmup + steps):
            t0 = time.perf_counter(); _cpu_worker(jobs[0]); dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
    else:
        with mp.get_context('fork').Pool(cores) as pool:
            for it in range(warmup + steps):
                t0 = time.perf_counter(); pool.map(_cpu_worker, jobs); dt = time.perf_counter() - t0
                if it >= warmup:
                    times.append(dt)
    sample = '%d window(s) of %s uint16, sigmas %s' % (cores, 'x'.join(map(str, jobs[0][0].shape)), jobs[0][1])
    return vox * len(times) / sum(times), float(np.mean(times)) * 1e3, sample


def reference_arm(args, rank):
    if rank != 0:
        return
    for k in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ.setdefault(k, '1')       # the reference is single-threaded; we scale by processes instead
    # a fixed core count (default 16, or all if the box has fewer) keeps the arm comparable between boxes
    cores = max(1, min(os.cpu_count() or 1, args.cpu_cores or 16))
    w = WORKLOADS[args.workload]
    # ~0.12 Mvox/s/core for the 3D path (BASELINE.md probe): ~0.5 Mvox per core per step is ~4-5 s
    v, ms, sample = run_cpu(args.workload, cores, args.steps, min(args.warmup, 1), args.cpu_voxels)
    line = {
        'impl': 'reference', 'metric': METRICS[args.workload], 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': args.workload, 'shape': list(w['shape']), 'sigmas': list(w['sig']), 'input_dtype': 'uint16',
                   'note': 'CPU arm: each step = one cropped window per core (bounded sample); warmup capped at 1; the '
                           'reference is single-threaded, the cores run independent windows (its own parfor suggestion, '
                           'calc_flow.py:512)', 'per_core_value': v / cores},
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': cpu_kind(), 'sample': sample},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def zslab_measure(rank, world, local_rank, sig, nz_total, ny, nx, precision, steps, warmup, chunk_planes, seed=1005, mode='dt'):
    """ONE output timepoint of a (kt, nz_total, ny, nx) uint16 window sharded by z-slab over the ranks of the default
    process group: every step = in-library NCCL halo exchange of the raw frames (straight into each rank's extended
    buffer, asynchronous) + of3d_flow3d_slab on the owned planes (chunks that do not touch a halo run during the
    exchange).  Returns the record on rank 0 (None elsewhere).  Also checks, untimed: (a) the received halo planes equal
    the planes the neighbour generated, (b) on a small volume, the sharded result is bit-identical to the single-GPU one."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from opticalflow3d_dev_b200 import _lib, multigpu
    from opticalflow3d_dev_b200.calc_flow import calc_flow3D
    dev = torch.device('cuda', local_rank)
    ctx = _lib.get_context(local_rank)

    def fill(zs_, z0, nzt):
        """this rank's owned planes of every frame, generated on the device"""
        own = zs_.own_frames()
        tmp = torch.empty((zs_.own, zs_.ny, zs_.nx), dtype=torch.uint16, device=dev)
        for k in range(zs_.kt):
            torch.cuda.synchronize()
            _lib.check(ctx.lib.of3d_synth_blobs(ctx.handle, tmp.data_ptr(), 1, zs_.own, zs_.ny, zs_.nx, k, z0, seed), 'synth')
            own[k].copy_(tmp)
        torch.cuda.synchronize()
        return tmp

    # ---- (b) parity on a small volume: sharded == unsharded, bit for bit
    H = multigpu.halo_planes(sig[0], sig[2])
    pn = (H + 7) * world
    small = multigpu.ZSlab(pn, 96, 128, np.uint16, sig, precision=precision, device=local_rank, chunk_planes=16, rel_dtype='float64',
                           exchange=mode)
    fill(small, small.z0, pn)
    small.exchange()
    part = small.flow()
    ctx.sync()
    parity = None
    gathered = [[torch.empty((multigpu.shard_timepoints(pn, world)[r][1] - multigpu.shard_timepoints(pn, world)[r][0], 96, 128),
                             dtype=part[0].dtype, device=dev) for r in range(world)] if rank == 0 else None for _ in range(4)]
    for i in range(4):
        dist.gather(part[i].contiguous(), gathered[i], dst=0)
    if rank == 0:
        full = torch.empty((small.kt, pn, 96, 128), dtype=torch.uint16, device=dev)
        torch.cuda.synchronize()
        _lib.check(ctx.lib.of3d_synth_blobs(ctx.handle, full.data_ptr(), small.kt, pn, 96, 128, 0, 0, seed), 'synth')
        ref = calc_flow3D(full, *sig, precision=precision, rel_dtype='float64', device=local_rank)
        parity = all(torch.equal(torch.cat(gathered[i], 0), ref[i]) for i in range(4))
        del full, ref
    small.close()
    del small, part, gathered
    torch.cuda.empty_cache()

    # ---- the measured volume
    zs_ = multigpu.ZSlab(nz_total, ny, nx, np.uint16, sig, precision=precision, device=local_rank, chunk_planes=chunk_planes,
                         rel_dtype='float64', exchange=mode)
    tmp = fill(zs_, zs_.z0, nz_total)
    outs = None
    chunk_used = zs_.chunk

    def step():
        zs_.exchange()
        return zs_.flow(out=outs)

    def barrier():
        ctx.sync(); torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()

    outs = list(step())
    ctx.sync()
    # ---- (a) the halo planes that arrived are the neighbour's planes (dt mode: the widened centre frame in `ic`)
    halo_ok = True
    kc = zs_.kt // 2
    dtm = mode in ('dt', 'dt_wide')

    def check_halo(z_first, count, got_raw, got_ic, t_raw):
        chk = torch.empty((count, ny, nx), dtype=torch.uint16, device=dev)
        torch.cuda.synchronize()
        _lib.check(ctx.lib.of3d_synth_blobs(ctx.handle, chk.data_ptr(), 1, count, ny, nx, kc if dtm else t_raw, z_first, seed), 'synth')
        return bool(torch.equal(chk.to(got_ic.dtype), got_ic)) if dtm else bool(torch.equal(chk, got_raw))

    if zs_.lo:
        halo_ok = halo_ok and check_halo(zs_.z0 - zs_.lo, zs_.lo, None if dtm else zs_.frames[zs_.kt - 1, :zs_.lo],
                                         zs_.ic[:zs_.lo] if dtm else None, zs_.kt - 1)
    if zs_.hi:
        halo_ok = halo_ok and check_halo(zs_.z1, zs_.hi, None if dtm else zs_.frames[0, zs_.lo + zs_.own:],
                                         zs_.ic[zs_.lo + zs_.own:] if dtm else None, 0)
    del tmp
    for _ in range(max(warmup - 1, 0)):
        step()
    barrier()
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    ctx.set_async(True)
    l0 = ctx.launch_count()
    ctx.stage_times()
    ctx.set_profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    ctx.set_async(False)
    barrier()
    ms = e0.elapsed_time(e1)
    stages = ctx.stage_times()
    ctx.set_profile(False)
    stat = torch.tensor([ms, float(ctx.launch_count() - l0), 0.0 if halo_ok else 1.0], dtype=torch.float64, device=dev)
    tmax = stat.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(stat, op=dist.ReduceOp.SUM)
    mem = torch.cuda.mem_get_info(local_rank)
    zs_.close()
    del zs_, outs
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    vol = nz_total * ny * nx
    value = vol * steps / (float(tmax[0].item()) * 1e-3)
    peak, peak_src = peaks()
    bpv = alg_bytes_per_voxel(sig, 3, precision)
    fma = alg_fma_per_voxel(sig, 3)
    fpk = fma_peak(precision)
    es = 2
    kt = 2 * math.ceil(3 * sig[1]) + 1
    return {
        'metric': 'output voxels/s (vx,vy,vz,rel) %dx%dx%d volume, z-slab sharded' % (nx, ny, nz_total), 'value': value, 'unit': UNIT,
        'n_gpus': world, 'steps': steps, 'ms_per_step': float(tmax[0].item()) / steps, 'scaling': 'weak' if nz_total != 512 else 'strong',
        'dtype': 'f64' if precision == 'fp64' else 'f32',
        'config': {'shape': [kt, nz_total, ny, nx], 'sigmas': list(sig), 'input_dtype': 'uint16', 'planes_per_rank': nz_total // world,
                   'chunk_planes': chunk_used,
                   'exchange': mode,
                   'sharding': ('z-slab; temporal stage on the boundary planes, then %d halo planes of (ic, dt0) exchanged with each '
                                'neighbour by the library (grouped ncclSend/ncclRecv, in place) while the interior temporal stage '
                                'runs; interior chunks overlap the exchange' % H) if mode in ('dt', 'dt_wide') else
                               ('z-slab; %d raw halo planes of each of the %d frames exchanged with each neighbour by the '
                                'library (grouped ncclSend/ncclRecv, in place); interior chunks overlap the exchange' % (H, kt)),
                   # 'dt': dt0 in the compute type + ic as the raw centre planes (8/16-bit frames); 'dt_wide': both in the compute type
                   'halo_bytes_sent_per_interior_rank_per_step': int(2 * H * ny * nx * (
                       ((8 if precision == 'fp64' else 4) + es) if mode == 'dt' else
                       (2 * (8 if precision == 'fp64' else 4) if mode == 'dt_wide' else kt * es)))},
        'parity_small_volume_bit_identical': parity, 'halo_planes_verified': bool(stat[2].item() == 0.0),
        'stages_ms_per_step_rank0': {k: round(v[0] / steps, 3) for k, v in stages.items()},
        'free_hbm_bytes_rank0_after': int(mem[0]),
        'roofline': {'bound': 'hbm', 'achieved': value * bpv / 1e9 / world, 'peak': peak, 'unit': 'GB/s',
                     'frac': value * bpv / 1e9 / world / peak, 'peak_source': peak_src, 'alg_bytes_per_voxel': bpv,
                     'fp_pipe': {'alg_fma_per_voxel': fma, 'achieved_tfma_per_s': value * fma / 1e12 / world,
                                 'peak_tfma_per_s': fpk, 'frac': value * fma / 1e12 / world / fpk}},
        'gpu_launches': int(stat[1].item()),
    }


def zslab_arm(args, rank, world, local_rank):
    """--workload cfg5: the 2048x2048x512 window of BASELINE config 5 over 2/4/8 GPUs"""
    import torch
    import torch.distributed as dist
    from opticalflow3d_dev_b200.build import build_library
    if rank == 0:
        build_library()
    if world < 2:
        raise SystemExit('workload cfg5 is z-slab sharded and needs --gpus >= 2')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    dist.barrier()
    torch.cuda.set_device(local_rank)
    w = WORKLOADS[args.workload]
    nt, nz, ny, nx = w['shape']
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    rec = zslab_measure(rank, world, local_rank, w['sig'], nz, ny, nx, args.precision, args.steps, args.warmup, args.chunk_planes,
                        mode=args.zslab_exchange)
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        line = dict(rec)
        line.update({'warmup': args.warmup, 'higher_is_better': True, 'vs_baseline': None, 'data': 'synthetic', 'clocks': clocks, 'e2e': None})
        line['config']['workload'] = args.workload
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


def device_measure(ctx, local_rank, rank, world, workload, precision, timepoints, steps, warmup, generic=False, stage_events=True,
                   sampler=None):
    """Device-timed pass of one workload: every output timepoint of the (possibly shortened) time-lapse once per step,
    sharded by output timepoint over the ranks, frames resident in HBM.  Returns a dict of raw measurements (this rank's)."""
    import ctypes as C
    import torch
    from opticalflow3d_dev_b200 import _lib
    from opticalflow3d_dev_b200.taps import flow_taps
    dev = torch.device('cuda', local_rank)
    lib = ctx.lib
    w = WORKLOADS[workload]
    shape, sig = tuple(w['shape']), w['sig']
    ndim = len(shape) - 1
    rt = math.ceil(3 * sig[1])
    kt = 2 * rt + 1
    nt = shape[0] if timepoints is None else min(shape[0], timepoints + 2 * rt)
    sp = shape[1:]
    nz = sp[0] if ndim == 3 else 1
    ny, nx = sp[-2], sp[-1]
    vol = int(np.prod(sp))
    outs_all = list(range(rt, nt - rt))                         # output timepoints of the time-lapse
    per = [len(outs_all) // world + (1 if r < len(outs_all) % world else 0) for r in range(world)]
    lo = sum(per[:rank]); mine = outs_all[lo:lo + per[rank]]
    taps, keep = _lib.make_taps(flow_taps(*sig))
    prec = _lib.FP64 if precision == 'fp64' else _lib.FP32
    odt = torch.float64 if precision == 'fp64' else torch.float32
    flags = _lib.FLAG_GENERIC if generic else 0

    # this rank's frames [mine[0]-rt, mine[-1]+rt], generated on the device (uint16 stored in int16 tensors)
    nloc = (len(mine) + 2 * rt) if mine else 0
    frames = torch.empty((max(nloc, 1), nz, ny, nx), dtype=torch.int16, device=dev)
    torch.cuda.synchronize()
    if mine:
        _lib.check(lib.of3d_synth_blobs(ctx.handle, frames.data_ptr(), nloc, nz, ny, nx, mine[0] - rt, 0, 1000 + int(workload[3:4])), 'synth')
    outs = [torch.empty((nz, ny, nx), dtype=odt, device=dev) for _ in range(ndim + 1)]
    optr = [C.c_void_p(o.data_ptr()) for o in outs]
    if ndim == 2:
        optr = [optr[0], optr[1], None, optr[2]]
    fbytes = vol * 2
    torch.cuda.synchronize()
    ctx.set_async(True)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    batch2d = 16 if (ndim == 2 and not generic and not os.environ.get('OF3D_BENCH_NO_BATCH2D')) else 0
    if batch2d:                                                  # 2D time-lapse: 16 output timepoints per launch set
        bouts = [torch.empty((batch2d, ny, nx), dtype=odt, device=dev) for _ in range(3)]

    def one_pass():
        if batch2d:
            for j0 in range(0, len(mine), batch2d):
                b = min(batch2d, len(mine) - j0)
                ptrs = (C.c_void_p * (b + kt - 1))(*[frames.data_ptr() + (j0 + k) * fbytes for k in range(b + kt - 1)])
                rc = lib.of3d_flow2d_batch(ctx.handle, ptrs, _lib.U16, b, ny, nx, C.byref(taps), prec, flags,
                                           *[C.c_void_p(o.data_ptr()) for o in bouts])
                _lib.check(rc, 'of3d_flow2d_batch')
            return
        for i in range(len(mine)):
            ptrs = (C.c_void_p * kt)(*[frames.data_ptr() + (i + k) * fbytes for k in range(kt)])
            rc = lib.of3d_flow_frames(ctx.handle, ndim, ptrs, _lib.U16, _lib.DEVICE, nz, ny, nx, C.byref(taps), prec, flags,
                                      optr[0], optr[1], optr[2], optr[3], _lib.DEVICE)
            _lib.check(rc, 'of3d_flow_frames')

    def barrier():
        ctx.sync(); torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier(); torch.cuda.synchronize()

    for _ in range(warmup):
        one_pass()
    barrier()
    if sampler is not None:
        sampler.start()
    l0 = ctx.launch_count()
    ctx.stage_times()                                            # clear the per-stage accumulators
    ctx.set_profile(stage_events)                                # event brackets around every launch, on the library stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        one_pass()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    stages = ctx.stage_times()
    ctx.set_profile(False)
    ctx.set_async(False)
    clocks = sampler.stop() if sampler is not None else None
    return dict(ms=ms, launches=launches, stages=stages, clocks=clocks, n_out=len(outs_all), n_mine=len(mine), vol=vol, sig=sig,
                ndim=ndim, nt=nt, sp=sp, kt=kt, rt=rt, frames=frames, nloc=nloc, batch2d=batch2d)


def roofline_record(m, workload, precision, value, world, steps, full):
    """`achieved` = the ALGORITHMIC bytes of the call (window read once, outputs written once:
    alg_bytes_per_voxel x the voxels one launch processes) / the dominant kernel's launch time -- against the measured
    copy peak.  What the kernel itself moves (its own inputs + outputs) is `kernel_local`; the roof that actually binds
    (CUDA-core FMA issue) is `fp_pipe`; every stage is in `stages`, the whole pipeline in `pipeline`."""
    peak, peak_src = peaks()
    sig, ndim, vol = m['sig'], m['ndim'], m['vol']
    bpv = alg_bytes_per_voxel(sig, ndim, precision)
    ach = value * bpv / 1e9 / world                             # per-GPU algorithmic GB/s of the whole pipeline
    fma = alg_fma_per_voxel(sig, ndim)
    fpk = fma_peak(precision)
    tr = measured_traffic(workload, precision) if full else None
    pipeline = {'alg_bytes_per_voxel': bpv, 'alg_gbs': ach, 'hbm_frac': ach / peak,
                'traffic_bytes_per_voxel': (tr['bytes_per_voxel'] if tr else None),
                'alg_fma_per_voxel': fma, 'tfma_per_s': value * fma / 1e12 / world,
                'fp_frac': (value * fma / 1e12 / world / fpk) if fpk else None,
                'launch': 'one output timepoint = %s kernels' % (tr['kernels'] if tr else 'several')}
    fused = 'temporal' not in m['stages']
    rows = stage_report(m['stages'], stage_model(sig, ndim, precision, fused=fused), vol, m['n_mine'] * steps, peak, fpk)
    top = next((r for r in rows if 'alg_gbs' in r), None)
    if not top:                                                  # generic kernels / stage events disabled
        return {'bound': 'hbm', 'kernel': 'whole pipeline', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                'traffic': (tr['bytes_per_voxel'] * vol if tr else None), 'peak_source': peak_src,
                'alg_bytes_per_voxel': bpv, 'alg_bytes_per_launch': bpv * vol, 'stages': rows, 'pipeline': pipeline}
    sd = (tr or {}).get('stage_dram_bytes_per_launch', {})
    k_ach = bpv * vol / top['ms_per_timepoint'] / 1e6
    return {'bound': 'hbm', 'kernel': top['stage'], 'achieved': k_ach, 'peak': peak, 'unit': 'GB/s', 'frac': k_ach / peak,
            'traffic': sd.get(top['stage']), 'peak_source': peak_src,
            'alg_bytes_per_voxel': bpv, 'alg_bytes_per_launch': bpv * vol,
            'launch_ms': top['ms_per_timepoint'], 'share_of_step': top['ms_per_timepoint'] * m['n_mine'] * steps / m['ms'],
            'note': 'achieved = algorithmic bytes of the call (SURVEY 8(d): %d B/voxel) x voxels per launch / the dominant '
                    'kernel\'s launch time' % bpv,
            'kernel_local': {'bytes_per_voxel': top['alg_bytes_per_voxel'], 'gbs': top['alg_gbs'], 'hbm_frac': top['hbm_frac'],
                             'note': 'the kernel\'s own inputs + outputs (incl. the intermediates it exchanges with its neighbours)'},
            'fp_pipe': {'alg_fma_per_voxel': top['alg_fma_per_voxel'], 'achieved_tfma_per_s': top['tfma_per_s'],
                        'peak_tfma_per_s': fpk, 'frac': top['fp_frac'],
                        'note': 'the roof that binds this kernel is the CUDA-core FMA pipe (measured peak, '
                                'tools/fma_peak.cu), not HBM; see DESIGN.md 3.3'},
            'stages': rows, 'pipeline': pipeline}


def gpu_arm(args, rank, world, local_rank):
    import torch
    from opticalflow3d_dev_b200.build import build_library
    if rank == 0:
        build_library()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        dist.barrier()
    from opticalflow3d_dev_b200 import _lib
    from opticalflow3d_dev_b200.calc_flow import calc_flow2D, calc_flow3D

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    from opticalflow3d_dev_b200 import numa
    cpus = numa.bind_to_device(local_rank) if world > 1 else None   # pinned buffers and copy threads on the GPU's NUMA node
    ctx = _lib.get_context(local_rank)
    m = device_measure(ctx, local_rank, rank, world, args.workload, args.precision, args.timepoints, args.steps, args.warmup,
                       generic=args.generic, stage_events=not args.no_stage_events, sampler=ClockSampler(local_rank) if rank == 0 else None)
    sig, ndim, vol, sp, kt, rt, nt = m['sig'], m['ndim'], m['vol'], m['sp'], m['kt'], m['rt'], m['nt']
    frames, nloc, clocks = m['frames'], m['nloc'], m['clocks']
    mine_n = m['n_mine']
    tmax = torch.tensor([m['ms']], dtype=torch.float64, device=dev)
    lsum = torch.tensor([float(m['launches'])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(lsum, op=dist.ReduceOp.SUM)
    ms_total = float(tmax.item())
    ms = m['ms']
    total_vox = m['n_out'] * vol * args.steps
    value = total_vox / (ms_total * 1e-3)

    def barrier():
        ctx.sync(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()

    # ---- end-to-end with HOST buffers, copies inside the timed region, two public entry points:
    #   stream : FlowStream (the engine under process_flow): every frame uploaded once from pinned memory, results
    #            copied back to pinned memory while the next window computes  -> the headline e2e
    #   call   : the synchronous drop-in calc_flow3D(host window) -> host arrays (re-uploads the whole window per call)
    e2e = None
    if not args.no_e2e:
        from opticalflow3d_dev_b200.timelapse import FlowStream
        n_e2e = max(1, min(args.e2e_timepoints, mine_n)) if mine_n else 0
        np_odt = np.float64 if args.precision == 'fp64' else np.float32
        call_rate = call_rate_pageable = None
        if n_e2e:
            nfr = n_e2e + 2 * rt + 2                             # frames fed to the stream: 2 warm-up windows + n_e2e timed
            nfr = min(nfr, nloc)
            hfr = _lib.pinned_empty((nfr,) + tuple(sp), np.uint16)
            hfr[...] = frames[:nfr].reshape(hfr.shape).cpu().numpy().view(np.uint16)
            eng = FlowStream(tuple(sp), np.uint16, sig, precision=args.precision, device=local_rank)
            n_warm = nfr - n_e2e                                 # pushes before the timed region (fills the ring + warm-up)
            for i in range(n_warm):
                eng.push(hfr[i], pinned=True)
            eng.flush()
        barrier()
        t0 = time.perf_counter()
        if n_e2e:
            b0, b1 = eng.h2d_bytes, eng.d2h_bytes
            for i in range(n_warm, nfr):
                eng.push(hfr[i], pinned=True)
            eng.flush()
        dt = time.perf_counter() - t0
        if n_e2e:
            h2d, d2h = eng.h2d_bytes - b0, eng.d2h_bytes - b1
            eng.close()
            del eng
        if world > 1:
            dist.barrier()                                       # the other ranks' streams are done: rank 0 measures alone
        if n_e2e and rank == 0:
            import gc
            gc.collect()
            # the synchronous per-window call, for comparison
            hout = tuple(_lib.pinned_empty(tuple(sp), np.float32 if (ndim == 3 and i == 3) else np_odt) for i in range(ndim + 1))
            fn = calc_flow3D if ndim == 3 else calc_flow2D
            kw = dict(precision=args.precision, device=local_rank, out=hout, generic=args.generic)
            fn(hfr[0:kt], *sig, **kw)
            tc = time.perf_counter()
            fn(hfr[1:1 + kt], *sig, **kw)
            call_rate = vol / (time.perf_counter() - tc)
            # the same call with ordinary (pageable) NumPy arrays in and out: what a script written for the reference does
            pg = np.array(hfr[0:kt + 1])
            kw2 = dict(precision=args.precision, device=local_rank, generic=args.generic)
            r_ = fn(pg[0:kt], *sig, **kw2); del r_
            tc = time.perf_counter()
            r_ = fn(pg[1:1 + kt], *sig, **kw2)
            call_rate_pageable = vol / (time.perf_counter() - tc)
            del r_, pg, hout
            gc.collect()
            _lib.pinned_pool_trim()
        if not n_e2e:
            h2d = d2h = 0
        # the box's ceiling for this number: every rank copies a pinned 1 GiB buffer device -> host for ~1 s, all ranks
        # at once (plain cudaMemcpyAsync, no engine); e2e moves d2h_bytes_per_voxel bytes per voxel over the same links
        ceil_gbs = 0.0
        if n_e2e or world > 1:
            cb_d = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
            cb_h = torch.from_numpy(_lib.pinned_empty((1 << 30,), np.uint8))
            cb_h.copy_(cb_d, non_blocking=True); torch.cuda.synchronize()
            barrier()
            tcp = time.perf_counter(); reps = 0
            while time.perf_counter() - tcp < 1.0:
                cb_h.copy_(cb_d, non_blocking=True); torch.cuda.synchronize(); reps += 1
            ceil_gbs = reps * (1 << 30) / (time.perf_counter() - tcp) / 1e9
            del cb_d, cb_h
            _lib.pinned_pool_trim()
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        nn = torch.tensor([float(n_e2e), float(h2d), float(d2h), ceil_gbs], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX); dist.all_reduce(nn, op=dist.ReduceOp.SUM)
        e2e = {'value': float(nn[0].item()) * vol / float(tt.item()), 'unit': UNIT,
               'h2d_bytes_per_step': int(nn[1].item()), 'd2h_bytes_per_step': int(nn[2].item()),
               'timepoints': int(nn[0].item()),
               'd2h_ceiling_gbs': float(nn[3].item()),
               'd2h_gbs': (float(nn[2].item()) / float(tt.item()) / 1e9) if float(tt.item()) > 0 else None,
               'frac_of_d2h_ceiling': (float(nn[2].item()) / float(tt.item()) / 1e9 / float(nn[3].item())) if float(nn[3].item()) > 0 else None,
               'ceiling_note': 'aggregate pinned device->host copy rate of all ranks copying at once, measured here (the GPUs of '
                               'this box share PCIe uplinks: tools/pcie_bw.py --gpus N, profiles/r02_pcie_concurrent_4gpu.jsonl)',
               'api': 'timelapse.FlowStream.push(host frame) -> host (vx,vy,vz,rel), pinned; the engine under process_flow',
               'numa_bound_cpus': (len(cpus) if cpus else None),
               'calc_flow_call_value': call_rate,
               'calc_flow_call_api': 'calc_flow%dD(host ndarray window) -> host ndarrays (pinned), rank 0' % ndim,
               'calc_flow_call_plain_numpy_value': call_rate_pageable}
        if n_e2e:
            del hfr
    del frames
    m['frames'] = None
    torch.cuda.empty_cache()

    # ---- z-slab sharding (BASELINE config 5), scaled to the ranks present: 64 owned planes of 2048x2048 per rank -- the
    # slab of config 5 at 8 GPUs, where this IS config 5 -- so that every multi-GPU line carries a measured z-slab record
    zrec = None
    if world > 1 and not args.no_zslab and args.workload == 'cfg4':
        w5 = WORKLOADS['cfg5']
        try:
            zrec = zslab_measure(rank, world, local_rank, w5['sig'], 64 * world, w5['shape'][2], w5['shape'][3], args.precision, 2, 1,
                                 args.chunk_planes, mode=args.zslab_exchange)
        except Exception as exc:                                  # the headline line must survive a failing sub-record
            zrec = {'error': '%s: %s' % (type(exc).__name__, exc)}

    # ---- the other BASELINE configs, device-timed on a few timepoints each (single-GPU runs)
    configs = None
    if world == 1 and not args.no_configs and args.workload == 'cfg4' and args.timepoints is None and not args.generic:
        configs = []
        for wl, prc, tps, stp in (('cfg1', 'fp64', 1, 50), ('cfg2', 'fp64', 25, 3), ('cfg3', 'fp64', 6, 3), ('cfg3', 'fp32', 6, 3),
                                  ('cfg4', 'fp32', 4, 2)):
            mm = device_measure(ctx, local_rank, 0, 1, wl, prc, tps, stp, 3)
            mm['frames'] = None
            torch.cuda.empty_cache()
            v = mm['n_out'] * mm['vol'] * stp / (mm['ms'] * 1e-3)
            rr = roofline_record(mm, wl, prc, v, 1, stp, False)
            configs.append({'workload': wl, 'metric': METRICS[wl], 'dtype': 'f64' if prc == 'fp64' else 'f32', 'value': v, 'unit': UNIT,
                            'shape': [mm['nt']] + list(mm['sp']), 'sigmas': list(mm['sig']), 'timepoints_per_step': mm['n_out'],
                            'steps': stp, 'ms_per_timepoint': mm['ms'] / (stp * mm['n_out']),
                            'api': ('of3d_flow2d_batch, %d timepoints per launch set' % mm['batch2d']) if mm['batch2d'] else 'of3d_flow_frames',
                            'hbm_frac': rr['pipeline']['hbm_frac'], 'fp_frac': rr['pipeline']['fp_frac'],
                            'alg_bytes_per_voxel': rr['pipeline']['alg_bytes_per_voxel'], 'alg_fma_per_voxel': rr['pipeline']['alg_fma_per_voxel'],
                            'stages_ms': {r['stage']: round(r['ms_per_timepoint'], 4) for r in rr['stages']}})

    if rank != 0:
        return
    full = args.shape is None and args.timepoints is None and not args.generic
    roof = roofline_record(m, args.workload, args.precision, value, world, args.steps, full)
    line = {
        'metric': METRICS[args.workload], 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f64' if args.precision == 'fp64' else 'f32', 'data': 'synthetic',
        'config': {'workload': args.workload, 'shape': [nt] + list(sp), 'sigmas': list(sig), 'input_dtype': 'uint16',
                   'output_timepoints_per_step': m['n_out'], 'sharding': 'output timepoint, no collective',
                   'l2': 'inputs+intermediates per timepoint (>= %.1f GB) exceed the 126 MB L2' % (kt * vol * 2 / 1e9),
                   'kernels': 'generic' if args.generic else 'default'},
        'roofline': roof,
        'clocks': clocks, 'gpu_launches': int(lsum.item()), 'e2e': e2e,
    }
    if zrec is not None:
        line['zslab'] = zrec
    if configs is not None:
        line['configs'] = configs
    if world == 1 and not args.no_cpu_baseline:
        v, _, sample = run_cpu(args.workload, 1, 1, 0, args.cpu_voxels * 4)
        line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': 1, 'kind': cpu_kind(), 'sample': sample}
    print(json.dumps(line), flush=True)


def _finish(world):
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg4', choices=sorted(WORKLOADS))
    ap.add_argument('--precision', default='fp64', choices=['fp64', 'fp32'])
    ap.add_argument('--timepoints', type=int, default=None, help='limit the number of output timepoints (debug)')
    ap.add_argument('--generic', action='store_true', help='force the generic kernels')
    ap.add_argument('--e2e-timepoints', type=int, default=12)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-stage-events', action='store_true', help='do not bracket the launches with CUDA events')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-voxels', type=int, default=400_000, help='voxels per core per CPU step')
    ap.add_argument('--cpu-cores', type=int, default=None, help='processes of the CPU arm (default 16, or all cores if fewer)')
    ap.add_argument('--chunk-planes', type=int, default=None, help='z-slab runs: owned planes per pass of the slab pipeline (default: as large as the free HBM allows)')
    ap.add_argument('--zslab-exchange', default='dt', choices=['dt', 'dt_wide', 'raw'],
                    help="halo exchange of (raw centre planes, dt0), of (ic, dt0) in the compute type, or of the raw frames")
    ap.add_argument('--no-zslab', action='store_true', help='multi-GPU runs: skip the z-slab sub-record')
    ap.add_argument('--no-configs', action='store_true', help='single-GPU runs: skip the quick lines of the other BASELINE configs')
    ap.add_argument('--shape', default=None, help='override the workload shape, e.g. 19,128,512,512 (debug)')
    args = ap.parse_args()
    if args.shape:
        WORKLOADS[args.workload] = dict(WORKLOADS[args.workload], shape=tuple(int(v) for v in args.shape.split(',')))
    rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        reference_arm(args, rank)
        return
    if args.gpus != world and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(args.gpus),
               '--master-addr', '127.0.0.1', '--master-port', str(29500 + os.getpid() % 1000)] + sys.argv
        os.execv(sys.executable, cmd)
    if WORKLOADS[args.workload].get('zslab'):
        zslab_arm(args, rank, world, local_rank)
        return
    gpu_arm(args, rank, world, local_rank)
    _finish(world)


if __name__ == '__main__':
    main()
