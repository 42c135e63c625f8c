"""
Pins the oracle: (1) against the committed golden fixtures (outputs of the unmodified
reference files, generated in the build container by tests/golden/make_golden.py),
everywhere; (2) against the reference files executed live, where /root/reference exists.
"""
import os

import numpy as np
import pytest

import golden_cases
from parity import rel_to_max, assert_structure
from oracle import ref_loader, stokes_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.mark.parametrize('name', sorted(golden_cases.CASES))
def test_oracle_matches_golden(name):
    func, args, kwargs = golden_cases.build(name)
    gold = np.load(os.path.join(GOLDEN, name + '.npz'))
    assert str(gold['func']) == func
    assert str(gold['input_sha256']) == golden_cases.digest(args, kwargs), \
        'synthetic inputs drifted from the ones the fixture was generated with'
    Y = getattr(stokes_oracle, func)(*args, **kwargs)
    assert_structure(Y.clm, Y.slm, int(gold['lmax']), int(gold['mmax']))
    # same arithmetic, same order: the restatement reproduces the reference to the bit,
    # allow a few ulp of the largest coefficient for libm differences between machines
    assert rel_to_max(Y.clm, Y.slm, gold['clm'], gold['slm']) < 1e-14


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
@pytest.mark.parametrize('name', ['pressure_options', 'atmosphere_weight', 'points_mmax_trunc'])
def test_oracle_matches_live_reference(name):
    ref = ref_loader.load()
    func, args, kwargs = golden_cases.build(name)
    A = getattr(ref, func)(*args, **kwargs)
    B = getattr(stokes_oracle, func)(*args, **kwargs)
    assert np.array_equal(A.clm, B.clm) and np.array_equal(A.slm, B.slm)
    assert int(A.lmax) == int(B.lmax) and int(A.mmax) == int(B.mmax)


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree not present')
def test_reference_grid_vs_point_consistency():
    """The reference's own test logic (test/test_point_pressures.py:16-76), seeded."""
    ref = ref_loader.load()
    from model_harmonics_b200 import testing as syn
    rng = np.random.default_rng(99)
    NPTS = 1500
    dlon = dlat = 1.0
    lat = np.arange(90.0 - dlat / 2.0, -90.0 - dlat / 2.0, -dlat)
    lon = np.arange(-180.0 + dlon / 2.0, 180.0 + dlon / 2.0, dlon)
    nlat, nlon = len(lat), len(lon)
    rad_e = np.atleast_1d(6.371000790009159e6)
    iy = rng.integers(0, nlat, NPTS)
    ix = rng.integers(0, nlon, NPTS)
    LAT, LON = lat[iy], lon[ix]
    th = np.radians(90.0 - LAT)
    PRESSURE = 100.0 - 200.0 * rng.standard_normal(NPTS)
    AREA = (rad_e**2) * np.sin(th) * np.radians(dlon) * np.radians(dlat)
    data = np.zeros((nlat, nlon))
    np.add.at(data, (iy, ix), PRESSURE)
    g = lambda t: 9.780356 * (1.0 + 5.2885e-3 * np.cos(t)**2 - 5.9e-6 * np.cos(2.0 * t)**2)
    gs = g(np.radians(90.0 - np.meshgrid(lon, lat)[1]))
    LOVE = syn.love_numbers(60)
    for impl in (ref, stokes_oracle):
        a = impl.gen_pressure_stokes(data, gs, rad_e, lon, lat, LMAX=60, LOVE=LOVE)
        b = impl.gen_point_pressure(PRESSURE, g(th), rad_e, LON, LAT, AREA=AREA, LMAX=60, LOVE=LOVE)
        eps = np.finfo(np.float32).eps
        assert np.all(np.abs(a.clm - b.clm) < eps)
        assert np.all(np.abs(a.amplitude - b.amplitude) < eps)
        # the two formulations agree to the disc-load approximation, not to rounding
        assert rel_to_max(a.clm, a.slm, b.clm, b.slm) < 1e-3


def test_oracle_error_floor_against_extended_precision():
    """The reference arithmetic itself sits ~1e-16 (relative to A_lm) from an 80-bit evaluation."""
    from oracle import error_scale as es
    func, args, kwargs = golden_cases.build('pressure_options')
    A = es.pressure_error_scale(*args, **kwargs)
    tc, ts = es.pressure_truth(*args, **kwargs)
    Y = stokes_oracle.gen_pressure_stokes(*args, **kwargs)
    ok = A > 0
    assert (np.abs(Y.clm - tc.astype(np.float64))[ok] / A[ok]).max() < 1e-14
    assert (np.abs(Y.slm - ts.astype(np.float64))[ok] / A[ok]).max() < 1e-14
