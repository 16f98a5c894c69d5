"""CPU tests: the oracle restatement against the reference's golden vectors (not gpu)."""
import numpy as np
import pytest

from conftest import GOLDEN_2D, GOLDEN_3D, load_golden
from oracle import lk_oracle as orc


@pytest.mark.parametrize('name', GOLDEN_3D)
def test_oracle_matches_reference_3d(name):
    g = load_golden(name)
    ss, ts, ws = g['sigmas']
    vx, vy, vz, rel = orc.lk_flow3d(g['images'], ss, ts, ws, rel_mode='reference')
    # flow: same operations in the same order as the reference -> bit-for-bit
    assert np.array_equal(vx, g['vx']) and np.array_equal(vy, g['vy']) and np.array_equal(vz, g['vz'])
    assert rel.dtype == np.float32
    assert np.array_equal(rel, g['rel'])
    # float64 reliability agrees with the reference's complex64 result to float32 resolution
    _, _, _, rel64, it = orc.lk_flow3d(g['images'], ss, ts, ws, rel_mode='float64', return_intermediates=True)
    lam_max = np.abs(it['wdx2']) + np.abs(it['wdy2']) + np.abs(it['wdz2'])
    assert np.all(np.abs(rel64 - g['rel']) <= 8 * np.finfo(np.float32).eps * lam_max)


@pytest.mark.parametrize('name', GOLDEN_2D)
def test_oracle_matches_reference_2d(name):
    g = load_golden(name)
    ss, ts, ws = g['sigmas']
    vx, vy, rel = orc.lk_flow2d(g['images'], ss, ts, ws)
    assert np.array_equal(vx, g['vx']) and np.array_equal(vy, g['vy'])
    assert np.array_equal(rel, g['rel'], equal_nan=True)


@pytest.mark.parametrize('name', ['g3_f', 'g2_e'])
def test_oracle_scipy_backend_identical(name):
    g = load_golden(name)
    ss, ts, ws = g['sigmas']
    if g['images'].ndim == 4:
        a = orc.lk_flow3d(g['images'], ss, ts, ws, rel_mode='reference', use_scipy=True)
    else:
        a = orc.lk_flow2d(g['images'], ss, ts, ws, use_scipy=True)
    for got, key in zip(a, ['vx', 'vy', 'vz', 'rel'] if len(a) == 4 else ['vx', 'vy', 'rel']):
        assert np.array_equal(got, g[key], equal_nan=True)


@pytest.mark.parametrize('taps', ['D', 'S', 'G', 'W', 'asym'])
@pytest.mark.parametrize('axis', [0, 1, 2])
def test_correlate1d_restatement_bit_exact_vs_scipy(taps, axis):
    from scipy.ndimage import correlate1d
    rng = np.random.default_rng(5)
    a = rng.normal(size=(7, 9, 11)) * 100
    if taps == 'asym':
        w = rng.normal(size=6)          # even length, no symmetry -> plain branch
    else:
        w = orc.make_taps(2.3, 1.5, 1.2)[taps]
    assert np.array_equal(orc.correlate1d_nearest(a, w, axis), correlate1d(a, w, axis=axis, mode='nearest'))


def test_taps_are_sampled_not_normalised():
    tp = orc.make_taps(1, 1, 4)
    assert [tp[k].size for k in 'DSGTW'] == [7, 3, 7, 7, 25]
    assert abs(tp['S'].sum() - 1.59684) < 1e-4          # un-normalised narrow smoother
    assert tp['D'][3] == 0.0 and np.array_equal(tp['D'], -tp['D'][::-1])
    tp = orc.make_taps(3, 3, 8)
    assert [tp[k].size for k in 'DSGTW'] == [19, 7, 19, 19, 49]
    tp = orc.make_taps(2.3, 1.5, 3.7)
    assert [tp[k].size for k in 'DSGTW'] == [15, 5, 15, 11, 25]


def test_oracle_input_errors():
    with pytest.raises(SystemExit, match='3D matrix with dimensions N_T, N_Z, N_Y, N_X'):
        orc.lk_flow3d(np.zeros((7, 8, 8)))
    with pytest.raises(SystemExit, match='3D matrix with dimensions N_T, N_Y, N_X'):
        orc.lk_flow2d(np.zeros((7, 4, 8, 8)))
    with pytest.raises(SystemExit, match='edge effects'):
        orc.lk_flow3d(np.zeros((5, 4, 8, 8)))
    with pytest.raises(SystemExit, match='odd number of timepoints'):
        orc.lk_flow2d(np.zeros((8, 8, 8)))


def test_oracle_constant_image_gives_zeros():
    vx, vy, vz, rel = orc.lk_flow3d(np.full((7, 6, 10, 12), 37, dtype=np.uint16), 1, 1, 2)
    assert not vx.any() and not vy.any() and not vz.any() and np.allclose(rel, 0)
