"""
Multi-GPU execution of the flow operator on one node; one process per GPU,
torch.distributed for the plumbing, the library's own NCCL communicator for the halo exchange.

* Timepoint sharding (time-lapses): output timepoint c depends only on frames c-Rt..c+Rt
  (calc_flow.py:276-278), so ranks take contiguous blocks of output timepoints and need NO collective.
  `shard_timepoints` computes the blocks; `timelapse.process_flow` and `bench.py` use it.

* z-slab sharding (volumes too large for one GPU): rank g owns planes [z0_g, z1_g).  The spatial stages reach
  R = ceil(3*xyzSig) planes (gradients) plus Rw = ceil(3*wSig) planes (window): H = R + Rw.  Each rank keeps every
  frame of the window in an EXTENDED buffer -- lo <= H halo planes, its own planes, hi <= H halo planes (none at the
  ends of the volume, where the reference's clamp-to-edge applies) -- and
      1. exchanges halo planes with its two neighbours, straight into place: one grouped ncclSend/ncclRecv on the
         library's exchange stream (of3d_halo_exchange), non-periodic.  exchange='dt' (default): the temporal stage
         (z-local) runs on the boundary planes first, the halos of (ic, dt0) travel while the temporal stage of the
         interior planes runs -- ic as the raw centre-frame planes it is the widening of for 8/16-bit integer frames
         (of3d_halo_exchange_centre: 2 + 8 bytes per voxel in fp64; exchange='dt_wide': 8 + 8, both in the compute
         type); exchange='raw': the raw planes of all kt frames (2 kt bytes per
         voxel: more bytes for kt > 8, but no temporal pass and no (ic, dt0) volumes);
      2. runs of3d_flow3d_slab(_dt): every stage works only on the planes the owned range needs (temporal derivative +
         gradient z pass and in-plane passes on own +- Rw, window sums and solve on own), in chunks of `chunk_planes`
         owned planes -- chunks that do not touch a halo run while the exchange is still in flight, and the
         workspace is bounded by the chunk, not by the slab.
  Every owned plane sees exactly the inputs it sees in the unsharded run, so the result is bit-identical to it.
  The exchange and compute functions are injectable so that the planning / exchange / cropping logic is testable on
  CPU ranks (gloo); the product path always uses the CUDA library.
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


# ---------------------------------------------------------------------------------------------- z-slab engine
def _np_dtype(t):
    return np.dtype(str(t.dtype).replace('torch.', ''))


def slab_workspace_bytes(chunk, plane, ts, rw, halo):
    """device workspace of one chunk of `chunk` owned planes (plan_bytes in csrc/of3d.cu, marching pipeline from (ic, dt0)
    or raw frames): the three z-pass volumes aliased with the nine z-windowed products, four gradient volumes on
    chunk + 2 Rw planes, and (ic, dt0) of chunk + 2 H planes when the temporal stage is not fused"""
    ng = chunk + 2 * rw
    return (max(3 * ng, 9 * chunk) + 4 * ng + 2 * (chunk + 2 * halo)) * plane * ts


def auto_chunk(own, plane, ts, rw, halo, budget):
    """chunk (<= own) of the smallest number of equal chunks whose workspace fits `budget` bytes, at least 8 planes
    (every chunk repeats the z marches over its 2 rw window planes: 64-plane chunks of a 49-tap window march 112 planes
    for 64, 86-plane chunks 134 for 86)"""
    for n in range(1, max(1, own // 8) + 1):
        c = -(-own // n)
        if c <= 8 or slab_workspace_bytes(c, plane, ts, rw, halo) <= budget:
            return int(max(c, min(own, 8)))
    return int(min(own, 8))


def exchange_frames_torch(ext, plan, rank, world, group=None):
    """Halo exchange of the extended frames tensor (kt, lo + own + hi, ...) with torch.distributed P2P (gloo on CPU
    ranks; the CUDA path uses the library's NCCL exchange).  Receives land in place: slices along z of every frame."""
    import torch.distributed as dist
    lo, hi = plan['lo'], plan['hi']
    own = ext.shape[1] - lo - hi
    plans = plan['_all']
    ops, tmp = [], []
    for k in range(ext.shape[0]):
        f = ext[k]
        if rank > 0:
            n = plans[rank - 1]['hi']                # what the neighbour below needs from us: ITS upper halo
            if n:
                ops.append(dist.P2POp(dist.isend, f[lo:lo + n].contiguous(), rank - 1, group))
            if lo:
                ops.append(dist.P2POp(dist.irecv, f[0:lo], rank - 1, group))
        if rank + 1 < world:
            n = plans[rank + 1]['lo']
            if n:
                ops.append(dist.P2POp(dist.isend, f[lo + own - n:lo + own].contiguous(), rank + 1, group))
            if hi:
                ops.append(dist.P2POp(dist.irecv, f[lo + own:lo + own + hi], rank + 1, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class ZSlab:
    """One rank's share of a z-slab sharded calc_flow3D: plan, extended frame buffer, exchange and flow.

        zs = ZSlab(nz_total, ny, nx, dtype, (xyzSig, tSig, wSig))           # needs an initialised process group
        zs.own_frames()[...] = my planes of the kt frames around the centre  # (kt, own, ny, nx) view, filled in place
        zs.exchange()                                                        # asynchronous
        vx, vy, vz, rel = zs.flow()                                          # owned planes, CUDA tensors
    """

    def __init__(self, nz_total, ny, nx, dtype, sigmas, *, rank=None, world=None, group=None, precision='fp64',
                 device=None, chunk_planes=None, rel_dtype='reference', use_nccl=True, exchange='dt'):
        import torch
        import torch.distributed as dist
        from . import _lib
        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.sig = tuple(sigmas)
        self.H = halo_planes(self.sig[0], self.sig[2])
        plans = plan_slabs(nz_total, self.world, self.H)
        self.plan = dict(plans[self.rank], _all=plans)
        self.z0, self.z1 = self.plan['own']
        self.lo, self.hi = self.plan['lo'], self.plan['hi']
        self.own = self.z1 - self.z0
        self.ny, self.nx = int(ny), int(nx)
        self.np_dt = np.dtype(dtype)
        self.code = _lib.DTYPE_CODES[self.np_dt]
        self.precision = precision
        self.dev = torch.cuda.current_device() if device is None else int(device)
        self.ctx = _lib.get_context(self.dev)
        self.taps, self._keep = _lib.make_taps(flow_taps(*self.sig))
        self.kt = self._keep[3].size
        tdev = torch.device('cuda', self.dev)
        tdt = getattr(torch, self.np_dt.name)
        if exchange not in ('dt', 'dt_wide', 'raw'):
            raise ValueError("exchange must be 'dt', 'dt_wide' or 'raw'")
        # 'dt': (ic, dt0) halos, ic as the raw centre planes for 8/16-bit integer frames; 'dt_wide': ic in the compute type
        self.raw_centre = exchange == 'dt' and self.np_dt in (np.dtype(np.uint8), np.dtype(np.uint16), np.dtype(np.int16))
        self._stage = None
        self.mode = 'dt' if exchange == 'dt_wide' else exchange
        self.odt = torch.float64 if precision == 'fp64' else torch.float32
        next_ = self.lo + self.own + self.hi
        if self.mode == 'raw':
            self.frames = torch.empty((self.kt, next_, self.ny, self.nx), dtype=tdt, device=tdev)
        else:
            self.frames = torch.empty((self.kt, self.own, self.ny, self.nx), dtype=tdt, device=tdev)
            self.ic = torch.empty((next_, self.ny, self.nx), dtype=self.odt, device=tdev)
            self.dt0 = torch.empty((next_, self.ny, self.nx), dtype=self.odt, device=tdev)
        self.rel_f32 = precision == 'fp64' and rel_dtype == 'reference'
        # chunk of owned planes per pass of the slab pipeline: bounds the workspace.  None: the largest chunk whose workspace
        # fits 70 % of the HBM that is free now (fewer chunks = less redundant gradient work on the 2 Rw planes around each)
        if chunk_planes is None:
            free = torch.cuda.mem_get_info(self.dev)[0]
            ts = 8 if precision == 'fp64' else 4
            outs = 4 * self.own * self.ny * self.nx * ts                       # flow() allocates the results afterwards
            self.chunk = auto_chunk(self.own, self.ny * self.nx, ts, math.ceil(3 * self.sig[2]), self.H, 0.7 * free - outs)
        else:
            self.chunk = int(chunk_planes)
        self._comm_ready = False
        self.use_nccl = use_nccl and self.world > 1

    def own_frames(self):
        """(kt, own, ny, nx) view to fill with this rank's planes of the kt frames around the centre"""
        return self.frames[:, self.lo:self.lo + self.own] if self.mode == 'raw' else self.frames

    def _temporal(self, p0, p1):
        """temporal stage of the owned planes [p0, p1) into the extended (ic, dt0) volumes (asynchronous)"""
        from . import _lib
        if p1 <= p0:
            return
        plane = self.ny * self.nx
        fb = self.own * plane * self.frames.element_size()
        off = p0 * plane * self.frames.element_size()
        ptrs = (C.c_void_p * self.kt)(*[self.frames.data_ptr() + k * fb + off for k in range(self.kt)])
        es = self.ic.element_size()
        o = (self.lo + p0) * plane * es
        rc = self.ctx.lib.of3d_temporal(self.ctx.handle, 3, ptrs, self.code, _lib.DEVICE, p1 - p0, self.ny, self.nx, C.byref(self.taps),
                                        _lib.FP64 if self.precision == 'fp64' else _lib.FP32, 0, self.ic.data_ptr() + o, self.dt0.data_ptr() + o)
        _lib.check(rc, 'of3d_temporal')

    def _init_comm(self):
        """rank 0's NCCL id reaches every rank through torch.distributed; the communicator lives in the library"""
        if self._comm_ready:
            return
        torch, dist = self.torch, self.dist
        from . import _lib
        buf = (C.c_char * 128)()
        if self.rank == 0:
            _lib.check(self.ctx.lib.of3d_comm_unique_id(buf), 'of3d_comm_unique_id')
        t = torch.frombuffer(bytearray(bytes(buf)), dtype=torch.uint8).clone()
        backend = dist.get_backend(self.group)
        if backend == 'nccl':
            t = t.cuda(self.dev)
        dist.broadcast(t, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
        raw = bytes(t.cpu().numpy().tobytes())
        idbuf = (C.c_char * 128).from_buffer_copy(raw)
        torch.cuda.synchronize(self.dev)
        _lib.check(self.ctx.lib.of3d_comm_init(self.ctx.handle, idbuf, self.world, self.rank), 'of3d_comm_init')
        self._comm_ready = True

    def exchange(self):
        """Start the halo exchange of every frame (returns at once; flow() waits for it on the device where needed).
        The caller's writes to own_frames() on torch's current stream are ordered before it."""
        from . import _lib
        torch = self.torch
        plans = self.plan['_all']
        send_dn = plans[self.rank - 1]['hi'] if self.rank > 0 else 0
        send_up = plans[self.rank + 1]['lo'] if self.rank + 1 < self.world else 0
        torch.cuda.current_stream(self.dev).synchronize()
        if self.mode == 'dt':
            # boundary planes first, then the exchange of (ic, dt0) starts while the interior planes are computed
            was_async = self.ctx.is_async
            self.ctx.set_async(True)
            try:
                if self.world == 1 or send_dn + send_up >= self.own:
                    self._temporal(0, self.own)
                    split = None
                else:
                    self._temporal(0, send_dn)
                    self._temporal(self.own - send_up, self.own)
                    split = (send_dn, self.own - send_up)
                if self.world > 1:
                    self._init_comm()
                    if self.raw_centre:
                        # ic travels as the raw centre planes it is the widening of (2 + 8 instead of 8 + 8 bytes per voxel)
                        if self._stage is None:
                            self._stage = torch.empty((max(1, self.lo + self.hi), self.ny, self.nx), dtype=self.frames.dtype,
                                                      device=self.frames.device)
                        rc = self.ctx.lib.of3d_halo_exchange_centre(
                            self.ctx.handle, self.frames[self.kt // 2].data_ptr(), self.code, self._stage.data_ptr(), self.ic.data_ptr(),
                            self.dt0.data_ptr(), _lib.FP64 if self.precision == 'fp64' else _lib.FP32, self.ny * self.nx, self.lo, self.own,
                            self.hi, send_dn, send_up)
                        _lib.check(rc, 'of3d_halo_exchange_centre')
                    else:
                        ptrs = (C.c_void_p * 2)(self.ic.data_ptr(), self.dt0.data_ptr())
                        rc = self.ctx.lib.of3d_halo_exchange(self.ctx.handle, ptrs, 2, self.ny * self.nx * self.ic.element_size(),
                                                             self.lo, self.own, self.hi, send_dn, send_up)
                        _lib.check(rc, 'of3d_halo_exchange')
                if split:
                    self._temporal(*split)
            finally:
                self.ctx.set_async(was_async)
            return
        if self.world == 1:
            return
        if not self.use_nccl:
            exchange_frames_torch(self.frames, self.plan, self.rank, self.world, self.group)
            return
        self._init_comm()
        fb = self.frames[0].numel() * self.frames.element_size()
        ptrs = (C.c_void_p * self.kt)(*[self.frames.data_ptr() + k * fb for k in range(self.kt)])
        rc = self.ctx.lib.of3d_halo_exchange(self.ctx.handle, ptrs, self.kt, self.ny * self.nx * self.frames.element_size(),
                                             self.lo, self.own, self.hi, send_dn, send_up)
        _lib.check(rc, 'of3d_halo_exchange')

    def flow(self, out=None):
        """(vx, vy, vz, rel) of the owned planes (CUDA tensors); the reliability is float32 with rel_dtype='reference'."""
        from . import _lib
        torch = self.torch
        tdev = self.frames.device
        odt = torch.float64 if self.precision == 'fp64' else torch.float32
        sp = (self.own, self.ny, self.nx)
        if out is None:
            out = [torch.empty(sp, dtype=odt, device=tdev) for _ in range(3)]
            out.append(torch.empty(sp, dtype=torch.float32 if self.rel_f32 else odt, device=tdev))
        torch.cuda.current_stream(self.dev).synchronize()
        flags = _lib.FLAG_REL_F32 if self.rel_f32 else 0
        if self.mode == 'dt':
            rc = self.ctx.lib.of3d_flow3d_slab_dt(self.ctx.handle, self.ic.data_ptr(), self.dt0.data_ptr(), self.lo + self.own + self.hi,
                                                  self.ny, self.nx, self.lo, self.own, self.chunk, C.byref(self.taps),
                                                  _lib.FP64 if self.precision == 'fp64' else _lib.FP32, flags,
                                                  *[o.data_ptr() for o in out])
            _lib.check(rc, 'of3d_flow3d_slab_dt')
            return tuple(out)
        fb = self.frames[0].numel() * self.frames.element_size()
        ptrs = (C.c_void_p * self.kt)(*[self.frames.data_ptr() + k * fb for k in range(self.kt)])
        rc = self.ctx.lib.of3d_flow3d_slab(self.ctx.handle, ptrs, self.code, self.lo + self.own + self.hi, self.ny, self.nx,
                                           self.lo, self.own, self.chunk, C.byref(self.taps),
                                           _lib.FP64 if self.precision == 'fp64' else _lib.FP32, flags,
                                           *[o.data_ptr() for o in out])
        _lib.check(rc, 'of3d_flow3d_slab')
        return tuple(out)

    def close(self):
        if self._comm_ready:
            self.ctx.lib.of3d_comm_destroy(self.ctx.handle)
            self._comm_ready = False


def calc_flow3D_zslab(frames_local, xyzSig=3, tSig=1, wSig=4, *, nz_total, rank=None, world=None, group=None, precision='fp64',
                      chunk_planes=None, rel_dtype='float64', exchange='dt', exchange_fn=None, flow_fn=None):
    """
    calc_flow3D on a volume sharded by z-slab.  `frames_local` is this rank's (Nt, nz_own, Ny, Nx) block (CUDA tensor;
    any strides), the blocks being the contiguous split of `nz_total` planes given by `plan_slabs`.  Nt follows the
    reference's rules (odd, >= 6*tSig+1, calc_flow.py:216-222); only the frames the temporal filter touches are used.
    Returns (vx, vy, vz, rel) for the planes this rank owns.  Requires an initialised process group.

    exchange_fn(ext, plan, rank, world, group) / flow_fn(ext, own_lo, own_n) replace the CUDA stages on CPU ranks (tests).
    """
    import sys
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group) if rank is None else rank
    world = dist.get_world_size(group) if world is None else world
    if frames_local.dim() != 4:
        sys.exit('ERROR: Input image must be a 3D matrix with dimensions N_T, N_Z, N_Y, N_X')
    nt = frames_local.shape[0]
    if nt < 6 * tSig + 1:
        sys.exit('ERROR: Input images will lead to edge effects. N_T must be >= 6*tSig+1')
    if not (nt % 2):
        sys.exit('ERROR: Input images must have an odd number of timepoints. Only the central time point is analyzed')
    kt = 2 * math.ceil(3 * tSig) + 1
    c0 = (nt + 1) // 2 - 1 - kt // 2                      # first frame the temporal filter touches (>= 0 by the checks)
    H = halo_planes(xyzSig, wSig)
    plans = plan_slabs(nz_total, world, H)
    plan = dict(plans[rank], _all=plans)
    z0, z1 = plan['own']
    if frames_local.shape[1] != z1 - z0:
        raise ValueError('rank %d owns planes [%d,%d) but got %d planes' % (rank, z0, z1, frames_local.shape[1]))
    ny, nx = frames_local.shape[2], frames_local.shape[3]
    lo, hi, own = plan['lo'], plan['hi'], z1 - z0
    if flow_fn is not None:                               # CPU ranks: same plan / exchange / crop logic, injected compute
        ext = torch.zeros((kt, lo + own + hi, ny, nx), dtype=frames_local.dtype)
        ext[:, lo:lo + own] = frames_local[c0:c0 + kt]
        (exchange_fn or exchange_frames_torch)(ext, plan, rank, world, group)
        return flow_fn(ext, lo, own)
    zs = ZSlab(nz_total, ny, nx, _np_dtype(frames_local), (xyzSig, tSig, wSig), rank=rank, world=world, group=group,
               precision=precision, device=frames_local.device.index, chunk_planes=chunk_planes, rel_dtype=rel_dtype,
               exchange=exchange)
    try:
        zs.own_frames().copy_(frames_local[c0:c0 + kt])   # strided views are fine: copy_ handles them
        zs.exchange()
        return zs.flow()
    finally:
        zs.close()


# ---------------------------------------------------------------------------------------------- two-stage entry points
def _cuda_temporal(frames, sig, precision, device):
    """of3d_temporal: frames, a CUDA tensor (Nt, nz, ny, nx) with any strides -> (ic, dt0) CUDA tensors of the compute
    type.  Nt follows the reference's rules (calc_flow.py:216-222)."""
    import sys
    import torch
    from . import _lib
    ctx = _lib.get_context(device)
    taps, keep = _lib.make_taps(flow_taps(*sig))
    kt = keep[3].size
    nt = frames.shape[0]
    if nt < 6 * sig[1] + 1:
        sys.exit('ERROR: Input images will lead to edge effects. N_T must be >= 6*tSig+1')
    if not (nt % 2):
        sys.exit('ERROR: Input images must have an odd number of timepoints. Only the central time point is analyzed')
    c0 = (nt + 1) // 2 - 1 - kt // 2
    win = frames[c0:c0 + kt].contiguous()               # frame pointers below assume a dense (kt, nz, ny, nx) block
    np_dt = _np_dtype(win)
    sp = tuple(win.shape[1:])
    odt = torch.float64 if precision == 'fp64' else torch.float32
    ic = torch.empty(sp, dtype=odt, device=win.device)
    dt0 = torch.empty(sp, dtype=odt, device=win.device)
    fb = win[0].numel() * win.element_size()
    ptrs = (C.c_void_p * kt)(*[win.data_ptr() + k * fb for k in range(kt)])
    torch.cuda.current_stream(win.device).synchronize()
    rc = ctx.lib.of3d_temporal(ctx.handle, 3, ptrs, _lib.DTYPE_CODES[np_dt], _lib.DEVICE, sp[0], sp[1], sp[2], C.byref(taps),
                               _lib.FP64 if precision == 'fp64' else _lib.FP32, 0, ic.data_ptr(), dt0.data_ptr())
    _lib.check(rc, 'of3d_temporal')
    ctx.sync()                                          # also when the shared context is in asynchronous mode
    return ic, dt0


def _cuda_spatial(ic, dt0, sig, precision, device):
    """of3d_flow_from_dt: (ic, dt0) CUDA tensors of the compute type -> [vx, vy, vz, rel]"""
    import torch
    from . import _lib
    ctx = _lib.get_context(device)
    taps, keep = _lib.make_taps(flow_taps(*sig))
    ic, dt0 = ic.contiguous(), dt0.contiguous()
    sp = tuple(ic.shape)
    outs = [torch.empty(sp, dtype=ic.dtype, device=ic.device) for _ in range(4)]
    torch.cuda.current_stream(ic.device).synchronize()
    rc = ctx.lib.of3d_flow_from_dt(ctx.handle, 3, ic.data_ptr(), dt0.data_ptr(), sp[0], sp[1], sp[2], C.byref(taps),
                                   _lib.FP64 if precision == 'fp64' else _lib.FP32, 0, *[o.data_ptr() for o in outs], _lib.DEVICE)
    _lib.check(rc, 'of3d_flow_from_dt')
    ctx.sync()
    return outs
