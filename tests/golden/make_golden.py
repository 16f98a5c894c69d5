"""
Generate golden vectors by running the UNMODIFIED reference here (build container only).

    python tests/golden/make_golden.py

Imports /root/reference/src/Python/calc_flow.py with stub modules for its two
I/O-only imports that are not installed (tifffile, natsort; calc_flow.py:12,16 --
neither is touched by calc_flow2D/calc_flow3D), runs calc_flow3D / calc_flow2D
on the seeded synthetic stacks listed in CASES, and stores input + outputs in
tests/golden/<name>.npz together with the library versions that produced them.
/root/reference does not exist on the GPU box; tests read only the .npz files.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from opticalflow3d_dev_b200.synth import make_stack  # noqa: E402

# name: (shape, (spatialSig, tSig, wSig), dtype, seed, generator kwargs)
CASES = {
    'g3_a': ((7, 9, 33, 41), (1, 1, 4), 'uint16', 11, {}),
    'g3_b': ((7, 12, 40, 37), (1.5, 1, 4), 'uint8', 12, dict(amp=(20, 120), noise=3.0)),
    'g3_c': ((7, 14, 40, 44), (3, 1, 4), 'float32', 13, {}),
    'g3_d': ((13, 17, 36, 37), (3, 2, 6), 'uint16', 14, {}),
    'g3_e': ((19, 10, 30, 28), (3, 3, 8), 'float64', 15, {}),
    'g3_f': ((11, 11, 29, 31), (2.3, 1.5, 3.7), 'int16', 16, {}),
    'g3_g': ((9, 8, 24, 26), (1, 1, 2), 'uint16', 17, {}),       # Nt larger than the tap count
    'g2_a': ((7, 65, 129), (1.5, 1, 4), 'uint16', 21, {}),
    'g2_b': ((7, 50, 47), (3, 1, 4), 'float32', 22, {}),
    'g2_c': ((13, 64, 80), (3, 2, 6), 'uint8', 23, dict(amp=(20, 120), noise=3.0)),
    'g2_d': ((19, 40, 33), (2, 3, 8), 'float64', 24, {}),
    'g2_e': ((11, 37, 45), (2.3, 1.5, 3.7), 'int16', 25, {}),
}


def import_reference():
    sys.modules.setdefault('tifffile', types.ModuleType('tifffile'))
    ns = types.ModuleType('natsort')
    ns.natsorted = sorted
    sys.modules.setdefault('natsort', ns)
    sys.path.insert(0, '/root/reference/src/Python')
    import calc_flow as ref
    assert ref.__file__.startswith('/root/reference/'), ref.__file__
    return ref


def main():
    import scipy
    ref = import_reference()
    for name, (shape, (ss, ts, ws), dtype, seed, kw) in CASES.items():
        img = make_stack(shape, seed=seed, dtype=np.dtype(dtype), **kw)
        keep = img.copy()
        if len(shape) == 4:
            vx, vy, vz, rel = ref.calc_flow3D(img, ss, ts, ws)
            out = dict(vx=vx, vy=vy, vz=vz, rel=np.ascontiguousarray(rel))
        else:
            vx, vy, rel = ref.calc_flow2D(img, ss, ts, ws)
            out = dict(vx=vx, vy=vy, rel=rel)
        assert np.array_equal(img, keep)
        np.savez_compressed(os.path.join(HERE, name + '.npz'), images=img,
                            sig=np.array([ss, ts, ws], dtype=np.float64),
                            versions=np.array([np.__version__, scipy.__version__]), **out)
        print(name, shape, dtype, {k: (v.dtype.name, float(np.nanmax(np.abs(v)))) for k, v in out.items()})


if __name__ == '__main__':
    main()
