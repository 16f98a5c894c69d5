"""
CPU oracle for the dense Lucas-Kanade hot path (TEST INFRASTRUCTURE ONLY).

This file is a NumPy restatement of what the reference computes in
``src/Python/calc_flow.py`` (``calc_flow2D`` lines 18-173, ``calc_flow3D`` lines
175-360).  It exists to *check* the CUDA path; nothing under
``opticalflow3d_dev_b200/`` imports it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline / ``--impl reference``)
may import it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified
reference in the build container and stores its outputs on seeded inputs;
``tests/test_oracle.py`` checks this restatement against those vectors
(flow fields bit-for-bit, 3D reliability to float32 resolution because the
reference runs LAPACK ``cgeev`` on complex64, ``calc_flow.py:355-357``).

Third-party arithmetic restated here (absent from /root/reference):
  * ``scipy.ndimage.correlate1d`` (pinned scipy 1.15.2 by
    ``src/Python/opticalflow3D.yml:200``): correlation (no kernel flip),
    origin 0, ``mode='nearest'`` = index clamp, and -- for odd-length taps that
    are symmetric or antisymmetric to DBL_EPSILON -- the paired summation
    ``out = a[c]*w[r];  out += (a[c+l] +/- a[c-l]) * w[r+l]  for l=-r..-1``.
    Restated in :func:`correlate1d_nearest`; ``tests/test_oracle.py`` checks it
    bit-for-bit against the installed scipy.
  * ``numpy.linalg.eigvals`` on a complex64 stack (LAPACK cgeev) for the 3D
    reliability.  The matrix is real symmetric, so the eigenvalues are real and
    the lexicographic complex minimum is the smallest real eigenvalue.
    ``rel_mode='reference'`` reproduces the call literally (float32 result);
    ``rel_mode='float64'`` uses a symmetric float64 solver and is the stricter
    target the CUDA fp64 mode is compared against.
"""
from __future__ import annotations

import math

import numpy as np

EPS = float(np.finfo(float).eps)  # calc_flow.py:155,338 -- additive regulariser

ERR_NDIM_2D = 'ERROR: Input image must be a 3D matrix with dimensions N_T, N_Y, N_X'      # calc_flow.py:55
ERR_NDIM_3D = 'ERROR: Input image must be a 3D matrix with dimensions N_T, N_Z, N_Y, N_X'  # calc_flow.py:213 (sic)
ERR_NT_SHORT = 'ERROR: Input images will lead to edge effects. N_T must be >= 6*tSig+1'    # calc_flow.py:61,219
ERR_NT_EVEN = ('ERROR: Input images must have an odd number of timepoints. '
               'Only the central time point is analyzed')                                  # calc_flow.py:64,222


class OracleInputError(SystemExit):
    """The reference calls sys.exit(msg) on bad input (calc_flow.py:54-64, 212-222)."""


# --------------------------------------------------------------------------- taps
def make_taps(spatial_sig, t_sig, w_sig):
    """Sampled (NOT normalised) filter taps, calc_flow.py:72-97 (2D) / 230-263 (3D).

    Returns dict with
      D : derivative-of-Gaussian, radius ceil(3*sig)           (fderiv*gderiv)
      S : narrow Gaussian of sigma sig/4, radius ceil(3*sig/4) (fsmooth)
      G : Gaussian of sigma sig, radius ceil(3*sig)            (fx)
      T : derivative-of-Gaussian in time, radius ceil(3*tSig)  (ft*gt)
      W : Lucas-Kanade window Gaussian, radius ceil(3*wSig)    (gw)
    The expressions keep the reference's operation order so the doubles are
    bit-identical to the reference's.
    """
    sig = spatial_sig
    x = np.arange(-math.ceil(3 * sig), math.ceil(3 * sig) + 1)
    sig2 = sig / 4
    y = np.arange(-math.ceil(3 * sig2), math.ceil(3 * sig2) + 1)
    fderiv = np.exp(-x * x / 2 / sig / sig) / math.sqrt(2 * math.pi) / sig
    fsmooth = np.exp(-y * y / 2 / sig2 / sig2) / math.sqrt(2 * math.pi) / sig2
    gderiv = x / sig / sig
    t = np.arange(-math.ceil(3 * t_sig), math.ceil(3 * t_sig) + 1)
    ft = np.exp(-t * t / 2 / t_sig / t_sig) / math.sqrt(2 * math.pi) / t_sig
    gt = t / t_sig / t_sig
    w = np.arange(-math.ceil(3 * w_sig), math.ceil(3 * w_sig) + 1)
    gw = np.exp(-w * w / 2 / w_sig / w_sig) / math.sqrt(2 * math.pi) / w_sig
    return {
        'D': np.ascontiguousarray(fderiv * gderiv, dtype=np.float64),
        'S': np.ascontiguousarray(fsmooth * 1, dtype=np.float64),
        'G': np.ascontiguousarray(fderiv * 1, dtype=np.float64),
        'T': np.ascontiguousarray(ft * gt, dtype=np.float64),
        'W': np.ascontiguousarray(gw, dtype=np.float64),
    }


# ------------------------------------------------------------------- correlate1d
def _symmetry(w):
    """+1 symmetric, -1 antisymmetric, 0 neither (scipy ni_filters.c NI_Correlate1D)."""
    n = w.size
    if not (n & 1):
        return 0
    r = n // 2
    c = w[r:]
    m = w[r::-1]
    de = np.finfo(np.float64).eps
    if np.all(np.abs(c[1:] - m[1:]) <= de):
        return 1
    if np.all(np.abs(c[1:] + m[1:]) <= de):
        return -1
    return 0


def correlate1d_nearest(a, w, axis):
    """Restatement of scipy.ndimage.correlate1d(a, w, axis=axis, mode='nearest').

    out[i] = sum_k w[k] * a[clamp(i + k - r)],  r = len(w)//2, float64 output,
    accumulated in scipy's order (paired for (anti)symmetric odd taps).
    """
    a = np.asarray(a, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    n = w.size
    r = n // 2            # size1
    r2 = n - r - 1        # size2
    a = np.moveaxis(a, axis, -1)
    L = a.shape[-1]
    pad = [(0, 0)] * (a.ndim - 1) + [(r, r2)]
    p = np.pad(a, pad, mode='edge')

    def sl(off):          # p[..., c+off] for every output position c
        return p[..., r + off: r + off + L]

    sym = _symmetry(w)
    if sym > 0:
        out = sl(0) * w[r]
        for l in range(-r, 0):
            out = out + (sl(l) + sl(-l)) * w[r + l]
    elif sym < 0:
        out = sl(0) * w[r]
        for l in range(-r, 0):
            out = out + (sl(l) - sl(-l)) * w[r + l]
    else:
        out = sl(r2) * w[r + r2]
        for l in range(-r, r2):
            out = out + sl(l) * w[r + l]
    return np.ascontiguousarray(np.moveaxis(out, -1, axis))


def _scipy_corr(a, w, axis):
    from scipy.ndimage import correlate1d
    return correlate1d(a, w, axis=axis, mode='nearest')


def _chain(a, filters, corr):
    """Apply [(taps, axis), ...] in order."""
    for w, ax in filters:
        a = corr(a, w, ax)
    return a


# ------------------------------------------------------------------ input checks
def check_inputs(shape, t_sig, ndim_expected):
    """calc_flow.py:54-65 (2D) / 212-223 (3D). Returns the centre-frame index."""
    if len(shape) != ndim_expected:
        raise OracleInputError(ERR_NDIM_2D if ndim_expected == 3 else ERR_NDIM_3D)
    nt = shape[0]
    if nt < 6 * t_sig + 1:
        raise OracleInputError(ERR_NT_SHORT)
    if not (nt % 2):
        raise OracleInputError(ERR_NT_EVEN)
    return math.ceil(nt / 2) - 1


# ----------------------------------------------------------------- reliability
def min_eig_sym3(xx, xy, xz, yy, yz, zz, mode='float64'):
    """Smallest eigenvalue of [[xx,xy,xz],[xy,yy,yz],[xz,yz,zz]] per voxel.

    mode='reference': literal calc_flow.py:352-357 (complex64 eigvals, lexicographic
    amin, real part) -> float32.
    mode='float64'  : numpy.linalg.eigvalsh in float64 (what MATLAB's pageeig on
    doubles computes, calc_flow3D.m:235-236) -> float64.
    """
    m = np.array([[xx, xy, xz], [xy, yy, yz], [xz, yz, zz]])
    m = np.moveaxis(m, [0, 1], [-1, -2])
    if mode == 'reference':
        ev = np.linalg.eigvals(m.astype(np.complex64))
        return np.real(np.amin(ev, axis=-1))
    ev = np.linalg.eigvalsh(m)
    return np.ascontiguousarray(ev[..., 0])


# ------------------------------------------------------------------------- 3D
def lk_flow3d(images, xyz_sig=3, t_sig=1, w_sig=4, rel_mode='float64', use_scipy=False,
              return_intermediates=False):
    """Restatement of calc_flow3D (calc_flow.py:175-360)."""
    c = check_inputs(images.shape, t_sig, 4)
    corr = _scipy_corr if use_scipy else correlate1d_nearest
    tp = make_taps(xyz_sig, t_sig, w_sig)
    D, S, G, T, W = tp['D'], tp['S'], tp['G'], tp['T'], tp['W']
    img = images.astype(np.float64)                                   # :225

    # :276-279  temporal derivative, keep the centre slice, then G in y, x, z
    rt = T.size // 2
    lo, hi = c - rt, c + rt + 1
    if lo >= 0 and hi <= img.shape[0]:
        # The centre slice only sees frames c-rt..c+rt; restricting the t-filter to
        # them is bit-identical to filtering all Nt frames (same operands, same order).
        dt0 = corr(img[lo:hi], T, 0)[rt]
    else:                                                             # unreachable given the checks
        dt0 = corr(img, T, 0)[c]
    ic = img[c]
    dt = _chain(dt0, [(G, 1), (G, 2), (G, 0)], corr)
    dy = _chain(ic, [(D, 1), (S, 2), (S, 0)], corr)                   # :282
    dx = _chain(ic, [(S, 1), (D, 2), (S, 0)], corr)                   # :285
    dz = _chain(ic, [(S, 1), (S, 2), (D, 0)], corr)                   # :288

    def win(p):                                                       # :300-313
        return _chain(p, [(W, 1), (W, 2), (W, 0)], corr)

    wdtx, wdty, wdtz = win(dx * dt), win(dy * dt), win(dz * dt)
    wdxy, wdxz, wdx2 = win(dx * dy), win(dx * dz), win(dx * dx)
    wdyz, wdy2, wdz2 = win(dy * dz), win(dy * dy), win(dz * dz)

    # :337-340
    det = (wdx2 * wdy2 * wdz2) + (2 * wdxy * wdxz * wdyz) - (wdy2 * wdxz ** 2) \
        - (wdz2 * wdxy ** 2) - (wdx2 * wdyz ** 2)
    inv = (det + EPS) ** -1
    vx = -inv * ((wdy2 * wdz2 - wdyz * wdyz) * wdtx + (wdxz * wdyz - wdxy * wdz2) * wdty
                 + (wdxy * wdyz - wdxz * wdy2) * wdtz)
    vy = -inv * ((wdyz * wdxz - wdxy * wdz2) * wdtx + (wdx2 * wdz2 - wdxz * wdxz) * wdty
                 + (wdxz * wdxy - wdx2 * wdyz) * wdtz)
    vz = -inv * ((wdxy * wdyz - wdy2 * wdxz) * wdtx + (wdxy * wdxz - wdx2 * wdyz) * wdty
                 + (wdx2 * wdy2 - wdxy * wdxy) * wdtz)
    rel = min_eig_sym3(wdx2, wdxy, wdxz, wdy2, wdyz, wdz2, rel_mode)   # :352-357
    if return_intermediates:
        inter = dict(dt0=dt0, dt=dt, dx=dx, dy=dy, dz=dz, wdtx=wdtx, wdty=wdty, wdtz=wdtz,
                     wdxy=wdxy, wdxz=wdxz, wdx2=wdx2, wdyz=wdyz, wdy2=wdy2, wdz2=wdz2)
        return vx, vy, vz, rel, inter
    return vx, vy, vz, rel


# ------------------------------------------------------------------------- 2D
def lk_flow2d(images, xy_sig=3, t_sig=1, w_sig=4, use_scipy=False, return_intermediates=False):
    """Restatement of calc_flow2D (calc_flow.py:18-173)."""
    c = check_inputs(images.shape, t_sig, 3)
    corr = _scipy_corr if use_scipy else correlate1d_nearest
    tp = make_taps(xy_sig, t_sig, w_sig)
    D, S, G, T, W = tp['D'], tp['S'], tp['G'], tp['T'], tp['W']
    img = images.astype(np.float64)                                   # :67

    rt = T.size // 2
    lo, hi = c - rt, c + rt + 1
    if lo >= 0 and hi <= img.shape[0]:
        dt0 = corr(img[lo:hi], T, 0)[rt]                              # :113-114
    else:
        dt0 = corr(img, T, 0)[c]
    ic = img[c]
    dt = _chain(dt0, [(G, 0), (G, 1)], corr)                          # :116
    dy = _chain(ic, [(D, 0), (S, 1)], corr)                           # :119
    dx = _chain(ic, [(S, 0), (D, 1)], corr)                           # :122

    def win(p):                                                       # :133-141
        return _chain(p, [(W, 0), (W, 1)], corr)

    wdtx, wdty = win(dx * dt), win(dy * dt)
    wdxy, wdx2, wdy2 = win(dx * dy), win(dx * dx), win(dy * dy)

    det = (wdx2 * wdy2) - (wdxy * wdxy)                               # :154
    inv = (det + EPS) ** -1
    vx = inv * ((wdy2 * -wdtx) + (-wdxy * -wdty))                     # :155
    vy = inv * ((-wdxy * -wdtx) + (wdx2 * -wdty))                     # :156
    trace = wdx2 + wdy2                                               # :163
    with np.errstate(invalid='ignore'):
        root = np.sqrt(trace ** 2 - 4 * det)                          # NaN if it rounds negative
    l1 = (trace + root) / 2
    l2 = (trace - root) / 2
    rel = np.real(np.minimum(l1, l2))                                 # :166-168
    if return_intermediates:
        inter = dict(dt0=dt0, dt=dt, dx=dx, dy=dy, wdtx=wdtx, wdty=wdty, wdxy=wdxy,
                     wdx2=wdx2, wdy2=wdy2)
        return vx, vy, rel, inter
    return vx, vy, rel
