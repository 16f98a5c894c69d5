"""File-to-file throughput of process_flow (the reference's entry point, calc_flow.py:362) on a RAM disk:
SequenceT TIFF frames in, vx/vy/vz/rel TIFFs out."""
import json, os, shutil, sys, time
import numpy as np
sys.path.insert(0, '.')
from opticalflow3d_dev_b200 import tiffio
from opticalflow3d_dev_b200.calc_flow import process_flow
from opticalflow3d_dev_b200.synth import make_stack

root = sys.argv[1] if len(sys.argv) > 1 else '/dev/shm/of3d_pf'
nt, sp = 16, (64, 512, 512)
shutil.rmtree(root, ignore_errors=True); os.makedirs(root)
stack = make_stack((nt,) + sp, seed=1, dtype=np.uint16)
for t in range(nt):
    tiffio.imwrite(os.path.join(root, 'exp_t%03d.tif' % t), stack[t])
out = {}
for rep in range(2):
    t0 = time.perf_counter()
    process_flow(root, 'exp_t.*', 'SequenceT', 3, 3, 1, 4, verbose=False)
    dt = time.perf_counter() - t0
    nout = nt - 6
    out['run%d' % rep] = {'s': dt, 'mvox_s': nout * int(np.prod(sp)) / dt / 1e6, 'timepoints': nout}
print(json.dumps(out))
shutil.rmtree(root, ignore_errors=True)
