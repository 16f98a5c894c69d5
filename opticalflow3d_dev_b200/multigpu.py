"""
Multi-GPU execution of the flow operator on one node; one process per GPU,
torch.distributed (NCCL over NVLink / NVSwitch) for the plumbing.

* Timepoint sharding (time-lapses): output timepoint c depends only on frames c-Rt..c+Rt
  (calc_flow.py:276-278), so ranks take contiguous blocks of output timepoints and need NO collective.
  `shard_timepoints` computes the blocks; `timelapse.process_flow` and `bench.py` use it.

* z-slab sharding (volumes too large for one GPU): rank g owns planes [z0_g, z1_g).  The temporal stage is
  local in z; the spatial stages reach R = ceil(3*xyzSig) planes (gradients) plus Rw = ceil(3*wSig) planes
  (window).  Each rank therefore
      1. runs of3d_temporal on its own planes                         -> ic, dt0  (device, compute type)
      2. exchanges H = R + Rw boundary planes of (ic, dt0) with its two neighbours: grouped NCCL
         send/recv, non-periodic (the global ends keep the reference's clamp-to-edge, they have no neighbour)
      3. runs of3d_flow_from_dt on the extended slab [z0-H, z1+H) n [0, Nz)
      4. keeps the planes it owns.  The planes within H of a cut are polluted by the clamp at the cut and are
         exactly the halo planes that step 4 drops, so the result is identical to the single-GPU result.
  The compute functions are injectable so that the sharding / exchange / cropping logic is testable on CPU
  ranks (gloo); the product path always uses the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from .taps import flow_taps


def shard_timepoints(n_out, world):
    """[(lo, hi)] per rank: contiguous blocks whose sizes differ by at most one."""
    base, extra = divmod(n_out, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def halo_planes(xyzSig, wSig):
    """z reach of the spatial stages: gradient radius + window radius (calc_flow.py:230, :262)."""
    return math.ceil(3 * xyzSig) + math.ceil(3 * wSig)


def plan_slabs(nz, world, halo):
    """Per rank: dict(own=(z0,z1), ext=(e0,e1), lo=planes received from below, hi=from above)."""
    own = shard_timepoints(nz, world)
    plans = []
    for r, (z0, z1) in enumerate(own):
        e0, e1 = max(0, z0 - halo), min(nz, z1 + halo)
        plans.append(dict(own=(z0, z1), ext=(e0, e1), lo=z0 - e0, hi=e1 - z1))
    for r, p in enumerate(plans):
        # the halo must come from the direct neighbour only (1-D, one exchange step)
        if (r > 0 and p['lo'] > own[r - 1][1] - own[r - 1][0]) or (r + 1 < world and p['hi'] > own[r + 1][1] - own[r + 1][0]):
            raise ValueError('z-slab sharding needs at least %d planes per rank (Nz=%d over %d ranks)' % (halo, nz, world))
    return plans


def exchange_halo(local, plan, rank, world, group=None):
    """local: tensor (nz_own, ...) of this rank.  Returns the extended tensor (lo + nz_own + hi, ...).
    One grouped send/recv with each neighbour (torch.distributed P2P: NCCL on GPUs, gloo on CPU ranks)."""
    import torch
    import torch.distributed as dist
    lo, hi = plan['lo'], plan['hi']
    parts, ops, recv_lo, recv_hi = [], [], None, None
    plans = plan['_all']      # what a neighbour needs from us is ITS halo width (differs from ours at clipped ends)
    if rank > 0:
        n_send = plans[rank - 1]['hi']
        if n_send:
            ops.append(dist.P2POp(dist.isend, local[:n_send].contiguous(), rank - 1, group))
        if lo:
            recv_lo = torch.empty((lo,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
            ops.append(dist.P2POp(dist.irecv, recv_lo, rank - 1, group))
    if rank + 1 < world:
        n_send = plans[rank + 1]['lo']
        if n_send:
            ops.append(dist.P2POp(dist.isend, local[local.shape[0] - n_send:].contiguous(), rank + 1, group))
        if hi:
            recv_hi = torch.empty((hi,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
            ops.append(dist.P2POp(dist.irecv, recv_hi, rank + 1, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if recv_lo is not None:
        parts.append(recv_lo)
    parts.append(local)
    if recv_hi is not None:
        parts.append(recv_hi)
    return torch.cat(parts, dim=0) if len(parts) > 1 else local


# ---------------------------------------------------------------------------------------------- CUDA stages
def _cuda_temporal(frames, sig, precision, device):
    """frames: CUDA tensor (Nt, nz, ny, nx) -> (ic, dt0) CUDA tensors of the compute type."""
    import torch
    from . import _lib
    ctx = _lib.get_context(device)
    taps, keep = _lib.make_taps(flow_taps(*sig))
    kt = keep[3].size
    nt = frames.shape[0]
    c0 = (nt + 1) // 2 - 1 - kt // 2
    np_dt = np.dtype(str(frames.dtype).replace('torch.', ''))
    sp = tuple(frames.shape[1:])
    odt = torch.float64 if precision == 'fp64' else torch.float32
    ic = torch.empty(sp, dtype=odt, device=frames.device)
    dt0 = torch.empty(sp, dtype=odt, device=frames.device)
    fb = frames[0].numel() * frames.element_size()
    ptrs = (C.c_void_p * kt)(*[frames.data_ptr() + (c0 + k) * fb for k in range(kt)])
    torch.cuda.current_stream(frames.device).synchronize()
    rc = ctx.lib.of3d_temporal(ctx.handle, 3, ptrs, _lib.DTYPE_CODES[np_dt], _lib.DEVICE, sp[0], sp[1], sp[2], C.byref(taps),
                               _lib.FP64 if precision == 'fp64' else _lib.FP32, 0, ic.data_ptr(), dt0.data_ptr())
    _lib.check(rc, 'of3d_temporal')
    return ic, dt0


def _cuda_spatial(ic, dt0, sig, precision, device):
    import torch
    from . import _lib
    ctx = _lib.get_context(device)
    taps, keep = _lib.make_taps(flow_taps(*sig))
    sp = tuple(ic.shape)
    outs = [torch.empty(sp, dtype=ic.dtype, device=ic.device) for _ in range(4)]
    torch.cuda.current_stream(ic.device).synchronize()
    rc = ctx.lib.of3d_flow_from_dt(ctx.handle, 3, ic.data_ptr(), dt0.data_ptr(), sp[0], sp[1], sp[2], C.byref(taps),
                                   _lib.FP64 if precision == 'fp64' else _lib.FP32, 0, *[o.data_ptr() for o in outs], _lib.DEVICE)
    _lib.check(rc, 'of3d_flow_from_dt')
    return outs


def calc_flow3D_zslab(frames_local, xyzSig=3, tSig=1, wSig=4, *, nz_total, rank=None, world=None, group=None, precision='fp64',
                      temporal_fn=None, spatial_fn=None):
    """
    calc_flow3D on a volume sharded by z-slab.  `frames_local` is this rank's (Nt, nz_own, Ny, Nx) block
    (CUDA tensor), the blocks being the contiguous split of `nz_total` planes given by `plan_slabs`.
    Returns (vx, vy, vz, rel) for the planes this rank owns.  Requires an initialised process group.
    """
    import torch.distributed as dist
    rank = dist.get_rank(group) if rank is None else rank
    world = dist.get_world_size(group) if world is None else world
    sig = (xyzSig, tSig, wSig)
    plans = plan_slabs(nz_total, world, halo_planes(xyzSig, wSig))
    plan = dict(plans[rank], _all=plans)
    z0, z1 = plan['own']
    if frames_local.shape[1] != z1 - z0:
        raise ValueError('rank %d owns planes [%d,%d) but got %d planes' % (rank, z0, z1, frames_local.shape[1]))
    dev = frames_local.device.index if getattr(frames_local, 'is_cuda', False) else None
    tfn = temporal_fn or (lambda fr: _cuda_temporal(fr, sig, precision, dev))
    sfn = spatial_fn or (lambda a, b: _cuda_spatial(a, b, sig, precision, dev))
    ic, dt0 = tfn(frames_local)                                   # stage 1, z-local
    ic_e = exchange_halo(ic, plan, rank, world, group)            # stage 2, neighbours only
    dt_e = exchange_halo(dt0, plan, rank, world, group)
    outs = sfn(ic_e, dt_e)                                        # stage 3 on the extended slab
    lo = plan['lo']
    return tuple(o[lo:lo + (z1 - z0)].contiguous() for o in outs)  # stage 4
