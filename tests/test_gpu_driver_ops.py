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
