"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/of3d.h declares, host-side argument checks mirror the reference, taps are bit-identical
to the oracle's.  No compute calls (no GPU here)."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from oracle import lk_oracle as orc


def _build():
    from opticalflow3d_dev_b200.build import build_library
    return build_library()


def test_library_builds_and_exports_every_declared_symbol():
    _build()
    from opticalflow3d_dev_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, 'include', 'of3d.h')).read()
    declared = set(re.findall(r'\b(of3d_[a-z0-9_]+)\s*\(', header))
    declared -= {'of3d_ctx', 'of3d_taps'}
    assert len(declared) >= 15
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.of3d_version() == 200


def test_no_cpu_fallback_without_device():
    _build()
    from opticalflow3d_dev_b200 import _lib
    lib = _lib.load()
    if lib.of3d_device_count() > 0:
        pytest.skip('a GPU is present')
    with pytest.raises(RuntimeError, match='no CUDA device'):
        _lib.Context(0)
    from opticalflow3d_dev_b200.calc_flow import calc_flow3D
    with pytest.raises(RuntimeError):
        calc_flow3D(np.zeros((7, 4, 8, 8), dtype=np.uint16))


@pytest.mark.parametrize('sig', [(1, 1, 4), (1.5, 1, 4), (3, 2, 6), (3, 3, 8), (2.3, 1.5, 3.7), (0.7, 0.5, 1.1)])
def test_host_taps_bit_identical_to_oracle(sig):
    from opticalflow3d_dev_b200.taps import flow_taps
    a, b = flow_taps(*sig), orc.make_taps(*sig)
    for k in 'DSGTW':
        assert a[k].dtype == np.float64 and np.array_equal(a[k], b[k]), k


def test_argument_checks_match_reference_messages():
    from opticalflow3d_dev_b200.calc_flow import calc_flow2D, calc_flow3D
    cases = [
        (calc_flow3D, np.zeros((7, 8, 8)), {}, orc.ERR_NDIM_3D),
        (calc_flow2D, np.zeros((7, 2, 8, 8)), {}, orc.ERR_NDIM_2D),
        (calc_flow3D, np.zeros((5, 2, 8, 8)), {}, orc.ERR_NT_SHORT),
        (calc_flow2D, np.zeros((8, 8, 8)), {}, orc.ERR_NT_EVEN),
        (calc_flow2D, np.zeros((9, 8, 8)), dict(tSig=1.5), orc.ERR_NT_SHORT),
        (calc_flow3D, np.zeros((12, 2, 8, 8)), dict(tSig=1.5), orc.ERR_NT_EVEN),
    ]
    for fn, img, kw, msg in cases:
        with pytest.raises(SystemExit) as e:
            fn(img, **kw)
        assert str(e.value) == msg
    # same messages as the oracle raises
    with pytest.raises(SystemExit) as e:
        orc.lk_flow3d(np.zeros((7, 8, 8)))
    assert str(e.value) == orc.ERR_NDIM_3D
