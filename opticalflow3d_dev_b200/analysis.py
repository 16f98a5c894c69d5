"""
Downstream analysis step of the reference (src/Python/example_analysis_script.ipynb cells 4-6), on the GPU:

    relThresh = np.percentile(rel, relPer);  relMask = rel > relThresh              (cell 4)
    v = v * relMask;  v[v == 0] = nan;  v = v * scale / tscale                      (cell 5)
    Magnitude, theta = arctan2(vy, vx), phi = arctan(vz / sqrt(vx^2 + vy^2))         (cell 6)

SURVEY.md section 8(f) rank 3: a fused epilogue saves reading the four result volumes back for users who only
want masked statistics.  Inputs are NumPy arrays (uploaded once) or torch CUDA tensors (e.g. straight from
calc_flow3D(cuda_tensor)); outputs live where the inputs live.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

__all__ = ['reliability_threshold', 'masked_flow']


def _is_cuda(x):
    return hasattr(x, 'data_ptr') and getattr(x, 'is_cuda', False)


def _to_device(x, device):
    import torch
    if _is_cuda(x):
        return x.contiguous()
    a = np.ascontiguousarray(x)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float64)
    return torch.from_numpy(a).to(torch.device('cuda', 0 if device is None else int(device)))


def percentile_plan(n, relPer, dtype):
    """The two 0-based ranks and the weight np.percentile(a, relPer) (method 'linear') interpolates with, for an array of
    n elements of float dtype `dtype`.  NumPy >= 2 carries the arithmetic out in the array's own dtype
    (numpy/lib/_function_base_impl.py: percentile -> _quantile -> _get_indexes / _get_gamma), which for float32
    reliabilities of more than 2**24 voxels quantises the virtual index; restated here so thresholds agree exactly."""
    dt = np.dtype(dtype).type
    q = dt(relPer) / dt(100)
    vi = dt(n - 1) * q
    if vi >= dt(n - 1):
        return n - 1, n - 1, dt(0)
    lo = int(np.floor(vi))
    return lo, min(lo + 1, n - 1), dt(vi - dt(lo))


def lerp(x, y, t):
    """numpy.lib._function_base_impl._lerp for scalars of one dtype: a + (b-a)*t, or b - (b-a)*(1-t) once t >= 0.5"""
    d = y - x
    return x + d * t if t < 0.5 else y - d * (1 - t)


def reliability_threshold(rel, relPer, device=None):
    """np.percentile(rel, relPer) (example_analysis_script.ipynb cell 4) by radix select on the GPU.
    Returns a NumPy scalar of rel's dtype, NaN if rel contains NaN (as NumPy does)."""
    import torch
    if not (0 <= relPer <= 100):
        raise ValueError('Percentiles must be in the range [0, 100]')
    t = _to_device(rel, device)
    f64 = t.dtype == torch.float64
    npdt = np.float64 if f64 else np.float32
    n = t.numel()
    lo, hi, gamma = percentile_plan(n, relPer, npdt)
    ctx = _lib.get_context(t.device.index)
    torch.cuda.current_stream(t.device).synchronize()
    a, b, nn = C.c_double(), C.c_double(), C.c_int64()
    _lib.check(ctx.lib.of3d_order_stats(ctx.handle, t.data_ptr(), int(f64), n, lo, hi, C.byref(a), C.byref(b), C.byref(nn)),
               'of3d_order_stats')
    if nn.value:
        return npdt(np.nan)
    with np.errstate(all='ignore'):
        return npdt(lerp(npdt(a.value), npdt(b.value), gamma))


def masked_flow(vx, vy, vz=None, rel=None, relPer=90, xyscale=1.0, zscale=1.0, tscale=1.0, device=None, relThresh=None):
    """Cells 4-6 in one call.  Returns dict(relThresh, vx, vy[, vz], Magnitude, theta[, phi]); velocities are masked
    (NaN where rel <= threshold or v == 0) and in physical units."""
    import torch
    if rel is None:
        raise ValueError('rel is required')
    on_device = _is_cuda(vx)
    tv = [_to_device(v, device) for v in ((vx, vy) if vz is None else (vx, vy, vz))]
    dev = tv[0].device
    tr = _to_device(rel, dev.index)
    if relThresh is None:
        relThresh = reliability_threshold(tr, relPer)
    if len({t.dtype for t in tv}) != 1 or any(t.shape != tr.shape for t in tv):
        raise ValueError('velocity components must share dtype and shape with rel')
    outs = [torch.empty_like(tv[0]) for _ in range(6 if vz is not None else 4)]
    ctx = _lib.get_context(dev.index)
    torch.cuda.current_stream(dev).synchronize()
    p = [o.data_ptr() for o in outs]
    if vz is None:
        ox, oy, mag, th = p
        args = (tv[0].data_ptr(), tv[1].data_ptr(), None, tr.data_ptr())
        oargs = (ox, oy, None, mag, th, None)
    else:
        ox, oy, oz, mag, th, ph = p
        args = (tv[0].data_ptr(), tv[1].data_ptr(), tv[2].data_ptr(), tr.data_ptr())
        oargs = (ox, oy, oz, mag, th, ph)
    _lib.check(ctx.lib.of3d_mask_derive(ctx.handle, *args, int(tv[0].dtype == torch.float64), int(tr.dtype == torch.float64),
                                        tr.numel(), float(relThresh), float(xyscale), float(zscale), float(tscale), *oargs),
               'of3d_mask_derive')
    names = ['vx', 'vy', 'Magnitude', 'theta'] if vz is None else ['vx', 'vy', 'vz', 'Magnitude', 'theta', 'phi']
    res = {'relThresh': relThresh}
    for nm, o in zip(names, outs):
        res[nm] = o if on_device else o.cpu().numpy()
    return res
