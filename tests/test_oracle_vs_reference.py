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


# ---------------------------------------------------------------------------------------------------
# Driver-side operations: the restatements against the drivers' OWN source lines,
# executed from /root/reference, and the radius-free sibling operator against gen_pressure_stokes.
# ---------------------------------------------------------------------------------------------------
def _driver_case():
    from model_harmonics_b200 import testing as syn
    nlat, nlon, T = 19, 36, 3
    lon, lat = syn.grid_axes(nlat, nlon)
    rng = np.random.default_rng(77)
    pressure = 1.0e5 + syn.pressure_fields(nlat, nlon, T, seed=70)
    ocean = rng.uniform(size=(nlat, nlon)) < 0.55
    ii, jj = np.nonzero(ocean)
    return lon, lat, pressure, pressure.mean(axis=0), ii, jj


@pytest.mark.skipif(not os.path.isdir('/root/reference'), reason='reference tree not present')
@pytest.mark.parametrize('weighted', [False, True])
def test_driver_ops_restatement_matches_the_driver_lines(weighted):
    from oracle import driver_ops_oracle as oo
    from model_harmonics_b200 import testing as syn
    lon, lat, pressure, mean_p, ii, jj = _driver_case()
    nlon = len(lon)
    weight = np.sin(np.radians(90.0 - lat)) * np.radians(10.0) if weighted else None
    M, N, gridtheta, dth, dphi = oo.grid_cell_area(lat, nlon)
    area = M * N * np.kron(np.ones((1, nlon)), weight[:, None]) if weighted else \
        (M * dth) * (N * np.sin(gridtheta) * dphi)
    for t in range(pressure.shape[0]):
        want = oo.reference_pressure_step(pressure, mean_p, t, lat, ii, jj, weight=weight)
        got = oo.redistribute_ocean_2d(oo.pressure_anomaly(pressure[t], mean_p), area, ii, jj)
        assert np.array_equal(got, want)
    GPH, dP, geoid = syn.atmosphere_cube(5, len(lat), nlon, seed=71)
    want = oo.reference_atmosphere_step(dP, lat, ii, jj, weight=weight)
    assert np.array_equal(oo.redistribute_ocean_levels(dP, area, ii, jj), want)


def _tile_case():
    rng = np.random.default_rng(78)
    Nt, nj, ni = 4, 9, 9
    lat = rng.uniform(-89.0, 89.0, (Nt, nj, ni))
    lon = rng.uniform(-180.0, 180.0, (Nt, nj, ni))
    obp = np.ma.array(50.0 * rng.standard_normal((Nt, nj, ni)), mask=rng.uniform(size=(Nt, nj, ni)) < 0.3)
    gamma_h = 9.8 + 0.02 * rng.standard_normal((Nt, nj, ni))
    R = 6.371e6 + 3.0e3 * rng.standard_normal((Nt, ni, nj))          # the driver indexes R[k, indi, indj]
    area = 1.0e10 * (1.0 + rng.uniform(size=(Nt, nj, ni)))
    return obp, gamma_h, R, area, lon, lat


@pytest.mark.skipif(not os.path.isdir('/root/reference'), reason='reference tree not present')
def test_tile_accumulation_of_the_driver_equals_one_concatenated_call():
    """ecco_llc_tile_harmonics.py:235-260 executed with the reference's gen_point_pressure, against one
    oracle call on the concatenated valid points (what gen_point_pressure_tiles does on the device)."""
    from oracle import driver_ops_oracle as oo, ref_loader, stokes_oracle
    from oracle.gravtk_standin import harmonics
    from model_harmonics_b200 import testing as syn
    obp, gamma_h, R, area, lon, lat = _tile_case()
    L = 20
    LOVE = syn.love_numbers(L)
    ref = ref_loader.load()
    want = oo.reference_tile_accumulation(obp, gamma_h, R, area, lon, lat, L, L, LOVE, ref.gen_point_pressure,
                                          harmonics)
    cols = [[], [], [], [], [], []]
    for k in range(obp.shape[0]):
        indj, indi = np.nonzero(~obp.mask[k])
        for c, a in zip(cols, (obp.data[k, indj, indi], gamma_h[k, indj, indi], R[k, indi, indj],
                               lon[k, indj, indi], lat[k, indj, indi], area[k, indj, indi])):
            c.append(a)
    P, G, Rr, lo, la, ar = (np.concatenate(c) for c in cols)
    got = stokes_oracle.gen_point_pressure(P, G, Rr, lo, la, AREA=ar, LMAX=L, MMAX=L, LOVE=LOVE)
    scale = np.abs(want.clm).max()
    assert np.abs(got.clm - want.clm).max() <= 1e-13 * scale and np.abs(got.slm - want.slm).max() <= 1e-13 * scale


def test_gen_stokes_restatement_equals_gen_pressure_stokes_without_radius():
    """gravtk.gen_stokes restated (unpinned third-party) == the pinned gen_pressure_stokes at R = rad_e, G = 1."""
    from oracle import stokes_oracle
    from oracle.gravtk_standin import units
    from model_harmonics_b200 import testing as syn
    L = 40
    LOVE = syn.love_numbers(L)
    lon, lat = syn.grid_axes(31, 60)
    data = syn.pressure_fields(31, 60, 1, seed=72)[0]
    rad_e = units(lmax=L).rad_e / 100.0
    want = stokes_oracle.gen_pressure_stokes(data, 1.0, rad_e, lon, lat, LMAX=L, MMAX=25, LOVE=LOVE)
    got = stokes_oracle.gen_stokes(data, lon, lat, LMAX=L, MMAX=25, UNITS=3, LOVE=LOVE)
    scale = np.abs(want.clm).max()
    assert np.abs(got.clm - want.clm).max() <= 1e-14 * scale and np.abs(got.slm - want.slm).max() <= 1e-14 * scale
    cm = stokes_oracle.gen_stokes(data.T, lon, lat, LMAX=L, MMAX=25, UNITS=1, LOVE=LOVE, LMIN=3)
    assert np.all(cm.clm[:3] == 0.0) and np.allclose(cm.clm[3:], 10.0 * got.clm[3:], rtol=1e-12, atol=0.0)
