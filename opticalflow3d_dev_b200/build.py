"""
Build libof3d.so in-tree with nvcc for sm_100a (B200).

    python -m opticalflow3d_dev_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, 'csrc')
LIB = os.path.join(PKG, 'libof3d.so')
SOURCES = ['of3d.cu', 'fast_f64.cu', 'fast_f32.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
              '--expt-relaxed-constexpr', '-Xptxas', '-v']


def _newest_source_mtime():
    m = 0.0
    for d in (CSRC, os.path.join(os.path.dirname(PKG), 'include')):
        for f in os.listdir(d):
            if f.endswith(('.cu', '.cuh', '.h', '.inc')):
                m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def needs_build():
    return (not os.path.exists(LIB)) or os.path.getmtime(LIB) < _newest_source_mtime()


def build_library(force=False, verbose=False):
    """Compile csrc/*.cu -> libof3d.so. Returns the library path."""
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objdir = os.path.join(PKG, 'build')
    os.makedirs(objdir, exist_ok=True)
    log = os.path.join(PKG, 'build.log')
    # one object per translation unit, compiled concurrently (the marching kernels are heavily unrolled)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get('OF3D_NVCC_EXTRA', '').split() + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((cmd, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    text, failed = '', False
    for cmd, obj, p in procs:
        out = p.communicate()[0]
        text += ' '.join(cmd) + '\n' + out
        failed |= p.returncode != 0
    if not failed:
        cmd = [nvcc, '-shared', '-o', LIB + '.tmp'] + [obj for _, obj, _ in procs] + ['-ldl']
        res = subprocess.run(cmd, capture_output=True, text=True)
        text += ' '.join(cmd) + '\n' + res.stdout + res.stderr
        failed = res.returncode != 0
    with open(log, 'w') as fh:
        fh.write(text)
    if failed:
        sys.stderr.write(text[-8000:])
        raise RuntimeError('nvcc failed (see %s)' % log)
    os.replace(LIB + '.tmp', LIB)
    if verbose:
        print(text)
    return LIB


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose=True))
