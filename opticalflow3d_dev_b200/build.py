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
SOURCES = ['of3d.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '-shared',
              '--expt-relaxed-constexpr', '-Xptxas', '-v']


def _newest_source_mtime():
    m = 0.0
    for d in (CSRC, os.path.join(os.path.dirname(PKG), 'include')):
        for f in os.listdir(d):
            if f.endswith(('.cu', '.cuh', '.h')):
                m = max(m, os.path.getmtime(os.path.join(d, f)))
    return m


def needs_build():
    return (not os.path.exists(LIB)) or os.path.getmtime(LIB) < _newest_source_mtime()


def build_library(force=False, verbose=False):
    """Compile csrc/*.cu -> libof3d.so. Returns the library path."""
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + [os.path.join(CSRC, s) for s in SOURCES] + ['-o', LIB + '.tmp']
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(PKG, 'build.log')
    with open(log, 'w') as fh:
        fh.write(' '.join(cmd) + '\n' + res.stdout + res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed (see %s)' % log)
    os.replace(LIB + '.tmp', LIB)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == '__main__':
    print(build_library(force='--force' in sys.argv, verbose=True))
