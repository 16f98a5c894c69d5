"""
Host-side filter taps, evaluated in float64 with the reference's own expressions so that
the doubles handed to the CUDA library are bit-identical to the reference's
(calc_flow.py:72-97 for 2D, :230-263 for 3D).

The taps are *sampled* Gaussians and are deliberately NOT normalised: e.g. the narrow
smoother at sigma = 1 sums to 1.597.  Reproducing that is part of drop-in parity.
"""
from __future__ import annotations

import math

import numpy as np


def _grid(sig):
    r = math.ceil(3 * sig)
    return np.arange(-r, r + 1)


def _gauss(x, sig):
    # operation order as in the reference: exp(-x*x/2/s/s)/sqrt(2*pi)/s
    return np.exp(-x * x / 2 / sig / sig) / math.sqrt(2 * math.pi) / sig


def flow_taps(spatialSig, tSig, wSig):
    """Return dict of float64 C-contiguous tap vectors D, S, G, T, W (see include/of3d.h)."""
    x = _grid(spatialSig)
    narrow = spatialSig / 4
    y = _grid(narrow)
    t = _grid(tSig)
    w = _grid(wSig)
    g = _gauss(x, spatialSig)
    taps = {
        'D': g * (x / spatialSig / spatialSig),   # derivative of Gaussian along the gradient axis
        'S': _gauss(y, narrow) * 1,               # narrow Gaussian (sigma/4) on the orthogonal axes
        'G': g * 1,                               # full Gaussian that smooths dI/dt
        'T': _gauss(t, tSig) * (t / tSig / tSig), # derivative of Gaussian in time
        'W': _gauss(w, wSig),                     # Lucas-Kanade neighbourhood weights
    }
    return {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in taps.items()}
