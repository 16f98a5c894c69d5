"""
Synthetic translating / deforming Gaussian-blob stacks (host, NumPy).

The reference ships no sample data (its inputs live on the authors' drive), so
tests and benchmarks use this seeded generator:

    I(t, p) = offset + sum_i A_i * exp(-|p - (c_i + u_i*(t - t_c))|^2 / (2 s_i(t)^2)) + noise

with sub-pixel velocities |u_i| <= 1 px/frame, a weak divergence field and slowly
breathing blob widths ("deforming"), additive Gaussian noise (keeps the structure
tensor away from singular, which makes relative-error metrics meaningful) and
camera-like quantisation to the requested dtype.

Large benchmark stacks are generated on the GPU by the library's
``of3d_synth_blobs`` kernel instead (same model, counter-based hashing).
"""
from __future__ import annotations

import numpy as np


def make_stack(shape, seed=0, dtype=np.uint16, offset=100.0, noise=5.0, density=1.0 / 16 ** 3,
               amp=(200.0, 1500.0), sigma=(2.0, 4.0), max_speed=1.0, divergence=0.002):
    """Return an array of ``shape`` = (Nt, Ny, Nx) or (Nt, Nz, Ny, Nx)."""
    shape = tuple(int(s) for s in shape)
    nt, sp = shape[0], shape[1:]
    nd = len(sp)
    if nd not in (2, 3):
        raise ValueError('shape must be (Nt,Ny,Nx) or (Nt,Nz,Ny,Nx)')
    rng = np.random.default_rng(seed)
    nvox = int(np.prod(sp))
    dens = density if nd == 3 else density ** (2.0 / 3.0)
    nb = max(4, int(round(nvox * dens)))
    ctr = rng.uniform(0, 1, size=(nb, nd)) * np.array(sp)
    vel = rng.normal(size=(nb, nd))
    vel *= (rng.uniform(0.2, max_speed, size=(nb, 1)) / np.linalg.norm(vel, axis=1, keepdims=True))
    vel += divergence * (ctr - np.array(sp) / 2.0)
    a = rng.uniform(*amp, size=nb)
    s0 = rng.uniform(*sigma, size=nb)
    ds = rng.uniform(-0.03, 0.03, size=nb)
    tc = (nt - 1) / 2.0
    out = np.full(shape, offset, dtype=np.float64)
    for t in range(nt):
        fr = out[t]
        for i in range(nb):
            s = s0[i] * (1.0 + ds[i] * (t - tc))
            c = ctr[i] + vel[i] * (t - tc)
            rad = int(np.ceil(4 * s))
            lo = [max(0, int(np.floor(c[d])) - rad) for d in range(nd)]
            hi = [min(sp[d], int(np.floor(c[d])) + rad + 1) for d in range(nd)]
            if any(h <= l for l, h in zip(lo, hi)):
                continue
            g = None
            for d in range(nd):
                x = np.arange(lo[d], hi[d]) - c[d]
                gd = np.exp(-x * x / (2 * s * s))
                g = gd if g is None else g[..., None] * gd
            fr[tuple(slice(l, h) for l, h in zip(lo, hi))] += a[i] * g
    if noise > 0:
        out += rng.normal(scale=noise, size=shape)
    dt = np.dtype(dtype)
    if dt.kind in 'ui':
        info = np.iinfo(dt)
        out = np.clip(np.rint(out), info.min, info.max)
    return out.astype(dt)
