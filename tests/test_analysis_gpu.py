"""Downstream analysis step against NumPy, following the reference's
example_analysis_script.ipynb cells 4-6 line by line."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _numpy_cells(vx, vy, vz, rel, relPer, xyscale, zscale, tscale):
    # example_analysis_script.ipynb cells 4-6, restated
    relThresh = np.percentile(rel, relPer)
    relMask = rel > relThresh
    vx = vx * relMask; vy = vy * relMask
    vx[vx == 0] = np.nan; vy[vy == 0] = np.nan
    vx = vx * xyscale / tscale; vy = vy * xyscale / tscale
    out = {'relThresh': relThresh, 'vx': vx, 'vy': vy}
    if vz is not None:
        vz = vz * relMask; vz[vz == 0] = np.nan; vz = vz * zscale / tscale
        out['vz'] = vz
        out['Magnitude'] = np.sqrt(np.power(vx, 2) + np.power(vy, 2) + np.power(vz, 2))
        out['phi'] = np.arctan(vz / np.sqrt(np.power(vx, 2) + np.power(vy, 2)))
    else:
        out['Magnitude'] = np.sqrt(np.power(vx, 2) + np.power(vy, 2))
    out['theta'] = np.arctan2(vy, vx)
    return out


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('n', [1, 2, 7, 1000, 1 << 20])
def test_percentile_matches_numpy(dtype, n):
    from opticalflow3d_dev_b200.analysis import reliability_threshold
    rng = np.random.default_rng(n)
    rel = (rng.standard_normal(n) * 10.0 ** rng.integers(-8, 8, n)).astype(dtype)
    if n > 100:
        rel[::17] = rel[3]          # ties
        rel[5] = 0.0; rel[6] = -0.0
    for q in (0, 1, 33.3, 50, 90, 99.9, 100):
        got = reliability_threshold(rel, q)
        want = np.percentile(rel, q)
        assert got.dtype == want.dtype
        assert got == want, (q, got, want)


def test_percentile_nan_and_range():
    from opticalflow3d_dev_b200.analysis import reliability_threshold
    rel = np.arange(100, dtype=np.float64)
    rel[10] = np.nan
    assert np.isnan(reliability_threshold(rel, 90))
    with pytest.raises(ValueError):
        reliability_threshold(rel, 101)


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
@pytest.mark.parametrize('ndim', [2, 3])
def test_masked_flow_matches_notebook_cells(dtype, ndim):
    from opticalflow3d_dev_b200.analysis import masked_flow
    rng = np.random.default_rng(5)
    shape = (9, 40, 50) if ndim == 3 else (70, 90)
    vx, vy, vz = (rng.standard_normal(shape).astype(dtype) for _ in range(3))
    vx[..., 3] = 0.0                                     # exact zeros become NaN even when reliable
    rel = rng.random(shape).astype(np.float32)
    if ndim == 2:
        vz = None
    want = _numpy_cells(vx.copy(), vy.copy(), None if vz is None else vz.copy(), rel, 90, 0.21, 0.5, 3.0)
    got = masked_flow(vx, vy, vz, rel, relPer=90, xyscale=0.21, zscale=0.5, tscale=3.0)
    assert got['relThresh'] == want['relThresh']
    for k in ('vx', 'vy') + (('vz',) if ndim == 3 else ()):
        assert got[k].dtype == want[k].dtype
        np.testing.assert_array_equal(got[k], want[k])           # mask, NaNs and scaling are bit-exact
    ulp = np.finfo(dtype).eps
    for k in ('Magnitude', 'theta') + (('phi',) if ndim == 3 else ()):
        assert np.array_equal(np.isnan(got[k]), np.isnan(want[k]))
        np.testing.assert_allclose(got[k], want[k], rtol=4 * ulp, atol=0, equal_nan=True)   # libm vs CUDA libm: <= 4 ulp


def test_masked_flow_on_calc_flow_output():
    """device-resident chain: calc_flow3D(cuda tensor) -> masked_flow, nothing returns to the host in between"""
    import torch
    from opticalflow3d_dev_b200 import calc_flow3D
    from opticalflow3d_dev_b200.analysis import masked_flow
    rng = np.random.default_rng(0)
    img = rng.integers(0, 4000, (7, 12, 48, 64)).astype(np.uint16)
    vx, vy, vz, rel = calc_flow3D(torch.from_numpy(img).cuda())
    res = masked_flow(vx, vy, vz, rel, relPer=75)
    assert res['vx'].is_cuda
    want = _numpy_cells(vx.cpu().numpy(), vy.cpu().numpy(), vz.cpu().numpy(), rel.cpu().numpy(), 75, 1.0, 1.0, 1.0)
    np.testing.assert_array_equal(res['vx'].cpu().numpy(), want['vx'])
    frac = np.isfinite(res['Magnitude'].cpu().numpy()).mean()
    assert 0.2 < frac < 0.26
