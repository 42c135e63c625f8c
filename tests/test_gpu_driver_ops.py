"""GPU: the drivers' per-step pre/post operations kept on the device."""
import numpy as np
import pytest

from parity import TOL, rel_to_max

pytestmark = pytest.mark.gpu


def test_anomaly_redistribution_and_mean_removal_chain():
    import torch
    import model_harmonics_b200 as mh
    from model_harmonics_b200 import driver_ops as ops, testing as syn
    from oracle import driver_ops_oracle as oo, stokes_oracle
    nlat, nlon, L, T = 46, 90, 30, 4
    lon, lat = syn.grid_axes(nlat, nlon)
    G, R = syn.surface_geometry(lon, lat, seed=2)
    rng = np.random.default_rng(5)
    pressure = 1.0e5 + syn.pressure_fields(nlat, nlon, T, seed=30)
    mean_p = pressure.mean(axis=0)
    ocean = rng.uniform(size=(nlat, nlon)) < 0.6
    ii, jj = np.nonzero(ocean)
    th = np.radians(90.0 - lat)
    area = (6.371e6 * np.radians(4.0)) * (6.371e6 * np.sin(th)[:, None] * np.radians(4.0)) * np.ones((1, nlon))
    LOVE = syn.love_numbers(L)
    # reference order of operations per time step
    want = []
    for t in range(T):
        P = oo.redistribute_ocean_2d(oo.pressure_anomaly(pressure[t], mean_p), area, ii, jj)
        want.append(stokes_oracle.gen_pressure_stokes(P, G, R, lon, lat, LMAX=L, LOVE=LOVE))
    mean_c = np.mean([y.clm for y in want], axis=0)
    mean_s = np.mean([y.slm for y in want], axis=0)
    # device chain: upload once, everything else on the GPU
    Pd = ops.redistribute_ocean(ops.pressure_anomaly(torch.from_numpy(pressure).cuda(), mean_p), area, ocean)
    assert np.abs(Pd[1].cpu().numpy() - oo.redistribute_ocean_2d(pressure[1] - mean_p, area, ii, jj)).max() < 1e-9
    clm, slm = mh.pressure_stokes_batch(Pd, G, R, lon, lat, LMAX=L, LOVE=LOVE, as_numpy=False)
    mc, ms = ops.mean_harmonics(clm, slm)
    c2, s2 = ops.subtract_mean_harmonics(clm, slm, mc, ms)
    for t in range(T):
        assert rel_to_max(clm[t].cpu().numpy(), slm[t].cpu().numpy(), want[t].clm, want[t].slm) <= TOL
        ref_c, ref_s = want[t].clm - mean_c, want[t].slm - mean_s
        scale = np.abs(want[t].clm).max()
        assert np.abs(c2[t].cpu().numpy() - ref_c).max() <= 1e-12 * scale
        assert np.abs(s2[t].cpu().numpy() - ref_s).max() <= 1e-12 * scale


def test_level_redistribution_matches_driver_loop():
    import torch
    from model_harmonics_b200 import driver_ops as ops, testing as syn
    from oracle import driver_ops_oracle as oo
    GPH, dP, geoid = syn.atmosphere_cube(6, 19, 36, seed=12)
    rng = np.random.default_rng(6)
    ocean = rng.uniform(size=(19, 36)) < 0.5
    ii, jj = np.nonzero(ocean)
    area = 1.0e9 * (1.0 + rng.uniform(size=(19, 36)))
    got = ops.redistribute_ocean(torch.from_numpy(dP).cuda(), area, ocean).cpu().numpy()
    want = oo.redistribute_ocean_levels(dP, area, ii, jj)
    assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    assert np.array_equal(got[:, ~ocean], dP[:, ~ocean])


def test_device_ops_match_the_drivers_own_lines_golden():
    """tests/golden/driver_ops.npz holds the outputs of the reference drivers' source lines
    (reanalysis_pressure_harmonics.py:474-490, reanalysis_atmospheric_harmonics.py:361-379) executed in the
    build container; the device forms must reproduce them to rounding."""
    import os
    import torch
    import golden_cases
    from model_harmonics_b200 import driver_ops as ops
    from oracle import driver_ops_oracle as oo
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'driver_ops.npz'))
    lon, lat, pressure, mean_p, ii, jj, dP, weight, tiles, L, LOVE = golden_cases.driver_case()
    nlon = len(lon)
    ocean = np.zeros(mean_p.shape, dtype=bool)
    ocean[ii, jj] = True
    M, N, gridtheta, dth, dphi = oo.grid_cell_area(lat, nlon)
    areas = {'plain': (M * dth) * (N * np.sin(gridtheta) * dphi),
             'weighted': M * N * np.kron(np.ones((1, nlon)), weight[:, None])}
    for name, area in areas.items():
        got = ops.redistribute_ocean(ops.pressure_anomaly(torch.from_numpy(pressure).cuda(), mean_p), area, ocean)
        want = gold['pressure_' + name]
        assert np.abs(got.cpu().numpy() - want).max() <= 1e-12 * np.abs(want).max()
        got = ops.redistribute_ocean(torch.from_numpy(dP).cuda(), area, ocean).cpu().numpy()
        want = gold['levels_' + name]
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()


def test_tile_accumulation_on_device_matches_the_driver_loop_golden():
    """gen_point_pressure_tiles (one kernel pass over the concatenated tiles) against the LLC driver's tile
    loop (ecco_llc_tile_harmonics.py:235-260) executed with the reference's gen_point_pressure."""
    import os
    import golden_cases
    import model_harmonics_b200 as mh
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'driver_ops.npz'))
    lon, lat, pressure, mean_p, ii, jj, dP, weight, tiles, L, LOVE = golden_cases.driver_case()
    obp, gamma_h, R, area, tlon, tlat = tiles
    tl = []
    for k in range(obp.shape[0]):
        indj, indi = np.nonzero(~obp.mask[k])
        tl.append((obp.data[k, indj, indi], gamma_h[k, indj, indi], R[k, indi, indj], tlon[k, indj, indi],
                   tlat[k, indj, indi], area[k, indj, indi]))
    Y = mh.gen_point_pressure_tiles(tl, LMAX=L, MMAX=L, LOVE=LOVE)
    assert rel_to_max(Y.clm, Y.slm, gold['tiles_clm'], gold['tiles_slm']) <= TOL
    # static geometry, three months at once
    months = [1.0, -0.5, 2.25]
    tb = [(np.stack([f * t[0] for f in months]),) + t[1:] for t in tl]
    clm, slm = mh.gen_point_pressure_tiles(tb, LMAX=L, MMAX=L, LOVE=LOVE)
    for i, f in enumerate(months):
        assert rel_to_max(clm[i], slm[i], f * gold['tiles_clm'], f * gold['tiles_slm']) <= TOL


def test_gen_stokes_sibling_operator():
    """gravtk.gen_stokes equivalent (radius factor = 1), all UNITS branches, both grid orientations."""
    import model_harmonics_b200 as mh
    from model_harmonics_b200 import testing as syn
    from oracle import stokes_oracle
    L = 60
    LOVE = syn.love_numbers(L)
    lon, lat = syn.grid_axes(91, 180)
    data = syn.pressure_fields(91, 180, 3, seed=73)
    th = np.radians(90.0 - lat)
    wgt = np.sin(th) * np.radians(2.0)**2 * 1.1
    cases = [dict(UNITS=1), dict(UNITS=3, MMAX=40), dict(UNITS=2), dict(UNITS=np.linspace(1.0, 2.0, L + 1)),
             dict(UNITS=3, WEIGHT=wgt, LMIN=2)]
    for kw in cases:
        Y = mh.gen_stokes(data[0], lon, lat, LMAX=L, LOVE=LOVE, **kw)
        O = stokes_oracle.gen_stokes(data[0], lon, lat, LMAX=L, LOVE=LOVE, **kw)
        assert rel_to_max(Y.clm, Y.slm, O.clm, O.slm) <= TOL, kw
    Y = mh.gen_stokes(data[0].T.copy(), lon, lat, LMAX=L, UNITS=3, LOVE=LOVE)
    O = stokes_oracle.gen_stokes(data[0], lon, lat, LMAX=L, UNITS=3, LOVE=LOVE)
    assert rel_to_max(Y.clm, Y.slm, O.clm, O.slm) <= TOL
    clm, slm = mh.stokes_batch(data, lon, lat, LMAX=L, UNITS=3, LOVE=LOVE)
    for t in range(3):
        O = stokes_oracle.gen_stokes(data[t], lon, lat, LMAX=L, UNITS=3, LOVE=LOVE)
        assert rel_to_max(clm[t], slm[t], O.clm, O.slm) <= TOL
    with pytest.raises(ValueError):
        mh.gen_stokes(data[0], lon, lat, LMAX=L, UNITS=9, LOVE=LOVE)
