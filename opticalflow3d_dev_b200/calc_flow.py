"""
calc_flow -- drop-in replacement for the reference module of the same name
(src/Python/calc_flow.py): dense Gaussian-weighted Lucas-Kanade optical flow for 2D+t and
3D+t image stacks, computed by hand-written CUDA kernels (sm_100a) through libof3d.so.

Public surface, mirrored from the reference:
    calc_flow2D(images, xySig=3, tSig=1, wSig=4)      reference calc_flow.py:18
    calc_flow3D(images, xyzSig=3, tSig=1, wSig=4)     reference calc_flow.py:175
    process_flow(imDir, imName, fileType="SequenceT", spatialDimensions=3,
                 xyzSig=3, tSig=1, wSig=4)            reference calc_flow.py:362

Same positional arguments, defaults, SystemExit messages and return conventions.  Extra
options are keyword-only and default to the reference's behaviour.  There is no CPU
fallback: without libof3d.so or without a CUDA device every call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

from . import _lib
from .taps import flow_taps

__all__ = ['calc_flow2D', 'calc_flow3D', 'calc_flow2D_timelapse', 'process_flow']

_MSG_NDIM_2D = 'ERROR: Input image must be a 3D matrix with dimensions N_T, N_Y, N_X'
_MSG_NDIM_3D = 'ERROR: Input image must be a 3D matrix with dimensions N_T, N_Z, N_Y, N_X'
_MSG_SHORT = 'ERROR: Input images will lead to edge effects. N_T must be >= 6*tSig+1'
_MSG_EVEN = 'ERROR: Input images must have an odd number of timepoints. Only the central time point is analyzed'


def _validate(images, tSig, ndim, msg_ndim):
    """The reference's argument checks, same order and text (calc_flow.py:54-64 / 212-222)."""
    if not (len(images.shape) == ndim):
        sys.exit(msg_ndim)
    Nt = images.shape[0]
    if Nt < 6 * tSig + 1:
        sys.exit(_MSG_SHORT)
    if not (Nt % 2):
        sys.exit(_MSG_EVEN)


def _is_cuda_tensor(x):
    return hasattr(x, 'data_ptr') and getattr(x, 'is_cuda', False)


def _default_device():
    env = os.environ.get('OF3D_DEVICE')
    if env is not None:
        return int(env)
    return 0


def _device_dtype(dt):
    """dtype a host array of dtype `dt` crosses the boundary in (the library widens it on the device)"""
    dt = np.dtype(dt)
    if dt.newbyteorder('=') in _lib.DTYPE_CODES:
        return dt.newbyteorder('=')
    if dt == np.bool_:
        return np.dtype(np.uint8)
    if dt == np.int8:
        return np.dtype(np.int16)
    if dt.kind in 'uif':
        return np.dtype(np.float64)       # what the reference does to every input (calc_flow.py:67,225)
    raise TypeError('calc_flow: unsupported image dtype %s' % dt)


def _upload_window(ctx, images, first, kt, chunk=0):
    """Upload the kt frames images[first:first+kt] into the library's device-resident window and return their dtype.
    Every piece must cross PCIe from PAGE-LOCKED memory: the driver's own staging of pageable memory moves a 1.9 GB
    window at a fifth of the PCIe rate.  Frames that already are pinned (e.g. _lib.pinned_empty), C-contiguous and of a
    supported dtype go straight; anything else is copied -- converting dtype / byte order / strides on the way -- into a
    pooled pinned block by several threads while earlier pieces are in flight (of3d_window_upload returns at once).
    chunk > 0 (3D): pieces are z-chunks of `chunk` planes shipped chunk-major -- every frame's chunk before the next
    chunk -- which is the order the library's z-slab pipeline consumes them in; otherwise whole frames."""
    a = np.asarray(images)
    dt = _device_dtype(a.dtype)
    win = a[first:first + kt]
    direct = win.dtype == dt and all(win[k].flags.c_contiguous for k in range(kt)) and _lib.is_pinned(win)
    stage = None
    if not direct:
        try:
            stage = _lib.pinned_empty(win.shape, dt, pooled=True)
        except RuntimeError:                                 # page-locking failed: let the driver stage pageable memory
            win, direct = np.ascontiguousarray(win, dtype=dt), True
    src = win if direct else stage
    nz = win.shape[1]
    fbytes = src[0].nbytes
    pbytes = fbytes // nz if nz else 0                       # bytes of one z plane (first spatial axis)
    if chunk > 0 and win.ndim == 4:
        cuts = [(z0, min(nz, z0 + chunk)) for z0 in range(0, nz, chunk)]
    else:
        cuts = [(0, nz)]
    pieces = [(k, z0, z1) for (z0, z1) in cuts for k in range(kt)]

    def ship(i):
        k, z0, z1 = pieces[i]
        _lib.check(ctx.lib.of3d_window_upload(ctx.handle, k, kt, C.c_void_p(src[k].ctypes.data + z0 * pbytes), fbytes,
                                              z0 * pbytes, (z1 - z0) * pbytes), 'of3d_window_upload')

    if direct or win.nbytes < (8 << 20):
        if not direct:
            np.copyto(stage, win, casting='unsafe')
        for i in range(len(pieces)):
            ship(i)
    else:
        _lib.parallel_copy_pieces([(stage[k, z0:z1], win[k, z0:z1]) for k, z0, z1 in pieces], ship)
    return dt, src                                           # keep the source alive until the flow call has synchronised


def _host_outputs(sp, dtypes):
    """Result arrays in pooled page-locked memory (plain NumPy arrays to the caller; the block goes back to the pool
    when the array dies).  Falls back to pageable memory when page-locking fails."""
    if os.environ.get('OF3D_PINNED_OUTPUTS', '1') != '0':
        try:
            return [_lib.pinned_empty(sp, d, pooled=True) for d in dtypes]
        except RuntimeError:
            pass
    return [np.empty(sp, dtype=d) for d in dtypes]


def _run(images, spatialSig, tSig, wSig, ndim, precision, device, exact, generic, out=None, rel_f32=False):
    if precision not in ('fp64', 'fp32'):
        raise ValueError("precision must be 'fp64' or 'fp32'")
    if not (spatialSig > 0 and tSig > 0 and wSig > 0):
        raise ValueError('sigmas must be positive')
    tp = flow_taps(spatialSig, tSig, wSig)
    taps, keep = _lib.make_taps(tp)
    prec = _lib.FP64 if precision == 'fp64' else _lib.FP32
    flags = (_lib.FLAG_EXACT if exact else 0) | (_lib.FLAG_GENERIC if generic else 0)
    if out is not None and precision == 'fp64':
        rel_f32 = getattr(out[-1], 'dtype', None) == np.float32      # the caller's buffer decides
    rel_f32 = bool(rel_f32) and precision == 'fp64'
    if rel_f32:
        flags |= _lib.FLAG_REL_F32                                   # float64 value, rounded once on the device
    nout = ndim + 1
    sp = tuple(int(s) for s in images.shape[1:])
    nt = int(images.shape[0])

    if _is_cuda_tensor(images):
        import torch
        t = images.contiguous()
        np_dt = np.dtype(str(t.dtype).replace('torch.', ''))
        if np_dt not in _lib.DTYPE_CODES:
            raise TypeError('calc_flow: unsupported tensor dtype %s' % t.dtype)
        dev = t.device.index if device is None else int(device)
        ctx = _lib.get_context(dev)
        odt = torch.float64 if precision == 'fp64' else torch.float32
        torch.cuda.current_stream(dev).synchronize()       # the library runs on its own stream
        outs = [torch.empty(sp, dtype=odt, device=t.device) for _ in range(nout)]
        if rel_f32:
            outs[-1] = torch.empty(sp, dtype=torch.float32, device=t.device)
        ptrs = [C.c_void_p(o.data_ptr()) for o in outs]
        in_ptr, code, mem = C.c_void_p(t.data_ptr()), _lib.DTYPE_CODES[np_dt], _lib.DEVICE
    else:
        ctx = _lib.get_context(_default_device() if device is None else device)
        kt = keep[3].size
        first = (nt + 1) // 2 - 1 - kt // 2                 # first frame the t-filter of the centre frame touches
        slab = int(ctx.lib.of3d_window_slab(ndim, sp[0] if ndim == 3 else 1, sp[-2], sp[-1], C.byref(taps)))
        in_dt, alive = _upload_window(ctx, images, first, kt, chunk=slab)
        odt = np.float64 if precision == 'fp64' else np.float32
        odts = [odt] * (nout - 1) + [np.float32 if rel_f32 else odt]
        if out is None:
            outs = _host_outputs(sp, odts)
        else:
            outs = list(out)
            if len(outs) != nout or any(not isinstance(o, np.ndarray) or o.shape != sp or o.dtype != d
                                        or not o.flags.c_contiguous for o, d in zip(outs, odts)):
                raise ValueError('out must be %d C-contiguous %s arrays of shape %s (the reliability may be float32)'
                                 % (nout, np.dtype(odt).name, sp))
        ptrs = [C.c_void_p(o.ctypes.data) for o in outs]
        o4 = ptrs if ndim == 3 else [ptrs[0], ptrs[1], None, ptrs[2]]
        rc = ctx.lib.of3d_window_flow(ctx.handle, ndim, _lib.DTYPE_CODES[in_dt], sp[0] if ndim == 3 else 1, sp[-2], sp[-1],
                                      C.byref(taps), prec, flags, o4[0], o4[1], o4[2], o4[3], _lib.HOST)
        _lib.check(rc, 'of3d_window_flow')
        del keep, alive
        return outs

    lib = ctx.lib
    if ndim == 3:
        rc = lib.of3d_flow3d(ctx.handle, in_ptr, code, mem, nt, sp[0], sp[1], sp[2], C.byref(taps), prec, flags,
                             ptrs[0], ptrs[1], ptrs[2], ptrs[3], mem)
    else:
        rc = lib.of3d_flow2d(ctx.handle, in_ptr, code, mem, nt, sp[0], sp[1], C.byref(taps), prec, flags,
                             ptrs[0], ptrs[1], ptrs[2], mem)
    _lib.check(rc, 'of3d_flow%dd' % ndim)
    del keep
    return outs


def calc_flow2D(images, xySig=3, tSig=1, wSig=4, *, precision='fp64', device=None, exact=False, generic=False,
                out=None):
    """
    Two-dimensional optical flow of the central frame of ``images`` (N_T, N_Y, N_X).

    Mirrors reference calc_flow2D (calc_flow.py:18-173): N_T must be odd and >= 6*tSig+1;
    (0,0) is the upper-left corner, so positive vy points down.

    Returns (vx, vy, rel): velocities in pixels/frame and the reliability (smallest
    eigenvalue of the windowed 2x2 structure tensor), each (N_Y, N_X), float64.

    Keyword-only extras (defaults reproduce the reference):
      precision  'fp64' (matches the reference to ~1e-12) or 'fp32' (float32 filters, float32 outputs)
      device     CUDA device index (default: $OF3D_DEVICE or 0; a CUDA tensor's own device)
      exact      use the bit-exact generic kernels (scipy's summation order, no FMA contraction)
      generic    force the generic kernels with normal rounding
      out        tuple of preallocated C-contiguous host arrays to receive the results; host inputs only
    Host results are ordinary NumPy arrays whose memory is page-locked and pooled (67 ms instead of ~700 ms to receive
    a 1024x1024x128 result; the block returns to the pool when the array dies; OF3D_PINNED_OUTPUTS=0 turns this off),
    and a pageable input window is staged into pinned memory by several threads before it is uploaded.
    A torch CUDA tensor may be passed instead of a NumPy array; outputs are then CUDA tensors.
    """
    _validate(images, tSig, 3, _MSG_NDIM_2D)
    vx, vy, rel = _run(images, xySig, tSig, wSig, 2, precision, device, exact, generic, out)
    return vx, vy, rel


def calc_flow3D(images, xyzSig=3, tSig=1, wSig=4, *, precision='fp64', device=None, exact=False, generic=False,
                rel_dtype='reference', out=None):
    """
    Three-dimensional optical flow of the central z-stack of ``images`` (N_T, N_Z, N_Y, N_X).

    Mirrors reference calc_flow3D (calc_flow.py:175-360).  Returns (vx, vy, vz, rel), each
    (N_Z, N_Y, N_X); velocities float64.

    rel_dtype='reference' (default) returns the reliability as float32, the dtype the reference
    returns because it runs the eigen-solver on complex64 (calc_flow.py:355-357); the value is
    computed in float64 and rounded once, on the device (OF3D_FLAG_REL_F32), so only 4 bytes per voxel
    cross PCIe.  With out=, the dtype of out[3] (float32 or float64) decides instead.  rel_dtype='float64' keeps the float64 value (what the
    MATLAB twin's pageeig on doubles gives, calc_flow3D.m:235-236).  Other keyword-only extras as
    in calc_flow2D.
    """
    _validate(images, tSig, 4, _MSG_NDIM_3D)
    if rel_dtype not in ('reference', 'float64'):
        raise ValueError("rel_dtype must be 'reference' or 'float64'")
    vx, vy, vz, rel = _run(images, xyzSig, tSig, wSig, 3, precision, device, exact, generic, out,
                           rel_f32=(rel_dtype == 'reference'))
    return vx, vy, vz, rel


def calc_flow2D_timelapse(images, xySig=3, tSig=1, wSig=4, *, precision='fp64', device=None, batch=16):
    """
    calc_flow2D of EVERY analysable frame of a 2D time-lapse ``images`` (N_T, N_Y, N_X): what the reference's loop over
    windows (calc_flow.py:599-606) computes one window at a time.  Returns (vx, vy, rel), each (N_T - 2R, N_Y, N_X) with
    R = ceil(3*tSig); row j belongs to frame j + R.  The frames are uploaded once and processed `batch` timepoints per
    launch set (of3d_flow2d_batch): identical values to per-window calc_flow2D calls, at about twice the rate on
    2048x2048 frames.  NumPy in -> NumPy out; CUDA tensor in -> CUDA tensors out.
    """
    if not (len(images.shape) == 3):
        sys.exit(_MSG_NDIM_2D)
    tp = flow_taps(xySig, tSig, wSig)
    taps, keep = _lib.make_taps(tp)
    kt = keep[3].size
    nt, ny, nx = (int(v) for v in images.shape)
    if nt < 6 * tSig + 1 or nt < kt:
        sys.exit(_MSG_SHORT)
    import torch
    cuda_in = _is_cuda_tensor(images)
    if cuda_in:
        dev = images.device.index if device is None else int(device)
        fr = images.contiguous()
    else:
        dev = _default_device() if device is None else int(device)
        a = np.ascontiguousarray(images, dtype=_device_dtype(np.asarray(images).dtype))
        fr = torch.from_numpy(a.view(np.int16) if a.dtype == np.uint16 else (a.view(np.int32) if a.dtype == np.uint32 else a)).cuda(dev)
        np_dt = a.dtype
    if cuda_in:
        np_dt = np.dtype(str(fr.dtype).replace('torch.', ''))
    if np_dt not in _lib.DTYPE_CODES:
        raise TypeError('calc_flow: unsupported image dtype %s' % np_dt)
    ctx = _lib.get_context(dev)
    n_out = nt - kt + 1
    odt = torch.float64 if precision == 'fp64' else torch.float32
    outs = [torch.empty((n_out, ny, nx), dtype=odt, device=fr.device) for _ in range(3)]
    fb = ny * nx * fr.element_size()
    plane_o = ny * nx * outs[0].element_size()
    torch.cuda.current_stream(dev).synchronize()
    bmax = max(1, min(int(batch), 129 - kt + 1))
    for j0 in range(0, n_out, bmax):
        b = min(bmax, n_out - j0)
        ptrs = (C.c_void_p * (b + kt - 1))(*[fr.data_ptr() + (j0 + i) * fb for i in range(b + kt - 1)])
        rc = ctx.lib.of3d_flow2d_batch(ctx.handle, ptrs, _lib.DTYPE_CODES[np_dt], b, ny, nx, C.byref(taps),
                                       _lib.FP64 if precision == 'fp64' else _lib.FP32, 0,
                                       *[C.c_void_p(o.data_ptr() + j0 * plane_o) for o in outs])
        _lib.check(rc, 'of3d_flow2d_batch')
    ctx.sync()
    if cuda_in:
        return tuple(outs)
    return tuple(o.cpu().numpy() for o in outs)


def process_flow(imDir, imName, fileType="SequenceT", spatialDimensions=3, xyzSig=3, tSig=1, wSig=4, **kwargs):
    """Time-lapse driver (reference calc_flow.py:362-625); implemented in .timelapse."""
    from .timelapse import process_flow as _pf
    return _pf(imDir, imName, fileType, spatialDimensions, xyzSig, tSig, wSig, **kwargs)
