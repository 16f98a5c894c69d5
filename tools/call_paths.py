"""Time the synchronous drop-in call calc_flow3D(host window) -> host arrays with pageable and with pinned buffers."""
import json, sys, time
import numpy as np
sys.path.insert(0, '.')
from opticalflow3d_dev_b200 import _lib, calc_flow3D
from opticalflow3d_dev_b200.synth import make_stack

sp = (128, 1024, 1024)
rng = np.random.default_rng(0)
img = rng.integers(0, 4000, (7,) + sp, dtype=np.uint16)
vol = int(np.prod(sp))
out = {}
calc_flow3D(img[:, :8], 3, 1, 4)                                   # load the library, warm the context
for name in ('pageable', 'pageable', 'pinned_out', 'pinned_in_out'):
    a = img
    kw = {}
    if name == 'pinned_in_out':
        a = _lib.pinned_empty(img.shape, img.dtype); a[...] = img
    if name != 'pageable':
        kw['out'] = tuple(_lib.pinned_empty(sp, np.float64) for _ in range(3)) + (_lib.pinned_empty(sp, np.float32),)
        calc_flow3D(a, 3, 1, 4, **kw)
    t0 = time.perf_counter()
    r = calc_flow3D(a, 3, 1, 4, **kw)
    dt = time.perf_counter() - t0
    out[name] = {'ms': dt * 1e3, 'gvox_s': vol / dt / 1e9}
    del r
print(json.dumps(out))
