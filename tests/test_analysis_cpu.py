"""Host logic of the analysis step: the rank/weight plan must reproduce np.percentile exactly (NumPy >= 2 semantics)."""
import numpy as np
import pytest

from opticalflow3d_dev_b200.analysis import lerp, percentile_plan


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('n', [1, 2, 3, 7, 100, 4097, 300001])
def test_percentile_plan_matches_numpy(dtype, n):
    rng = np.random.default_rng(n)
    a = (rng.standard_normal(n) * 10.0 ** rng.integers(-6, 6, n)).astype(dtype)
    s = np.sort(a)
    for q in (0, 1, 10, 33.3, 50, 75, 90, 95, 99.9, 100):
        lo, hi, g = percentile_plan(n, q, dtype)
        got = lerp(s[lo], s[hi], g)
        want = np.percentile(a, q)
        assert type(got) is type(want) and got == want, (q, lo, hi, g, got, want)


def test_percentile_plan_large_float32_index_quantisation():
    # beyond 2**24 elements NumPy's float32 virtual index is quantised; the plan must follow it
    n = (1 << 25) + 12345
    rng = np.random.default_rng(1)
    a = rng.random(n, dtype=np.float32)
    lo, hi, g = percentile_plan(n, 90, np.float32)
    part = np.partition(a, [lo, hi])
    assert lerp(part[lo], part[hi], g) == np.percentile(a, 90)


def test_threaded_host_copies():
    """_lib.parallel_copy / parallel_copy_frames (staging of pageable windows): values, dtype conversion, strides, order"""
    from opticalflow3d_dev_b200 import _lib
    rng = np.random.default_rng(0)
    a = rng.integers(0, 60000, (5, 16, 96, 130)).astype(np.uint16)
    b = np.empty(a.shape, np.float64)
    seen = []
    _lib.parallel_copy_frames(b, a, seen.append)
    assert seen == [0, 1, 2, 3, 4] and np.array_equal(b, a)
    c = np.empty((5, 16, 96, 65), np.uint16)
    _lib.parallel_copy_frames(c, a[..., ::2].astype('>u2'), lambda k: None)      # strided + byte-swapped source
    assert np.array_equal(c, a[..., ::2])
    d = np.empty_like(a)
    _lib.parallel_copy(d, a, min_bytes=1)
    assert np.array_equal(d, a)
    one = np.empty(7)
    _lib.parallel_copy_frames(one, np.arange(7.), lambda k: None)
    assert np.array_equal(one, np.arange(7.))
