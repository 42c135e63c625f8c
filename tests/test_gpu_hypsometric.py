"""GPU parity of the upstream producer and of the chained producer -> Stokes path."""
import os

import numpy as np
import pytest

from parity import TOL, rel_to_max

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'hypsometric.npz')


def _ulp_err(a, ref):
    return np.max(np.abs(a.astype(np.float64) - ref.astype(np.float64)) / np.spacing(np.abs(ref)))


@pytest.mark.parametrize('kw', [dict(nlev=11, nlat=10, nlon=18, seed=20), dict(nlev=37, nlat=46, nlon=90, seed=6),
                                dict(nlev=137, nlat=19, nlon=36, seed=8),
                                dict(nlev=7, nlat=5, nlon=8, seed=4, interfaces_extra=False)])
def test_kernel_matches_oracle(kw):
    import model_harmonics_b200 as mh
    from oracle import hypsometric_oracle as ho
    args = ho.synthetic_model_levels(**kw)
    Zo, Do = ho.geopotential_levels(*args)
    Z, D = mh.geopotential_levels(*args)
    assert Z.dtype == np.float32 and D.dtype == np.float32 and Z.shape == Zo.shape
    assert np.array_equal(D, Do)                       # no transcendental on this output: bit exact
    # log() is the only operation that is not bit-defined (<= 1 ulp in float64): at most one float32 ulp
    assert _ulp_err(Z, Zo) <= 1.0
    frac_exact = np.mean(Z == Zo)
    print('fraction of bit-identical geopotentials: %.6f' % frac_exact)
    assert frac_exact > 0.999


def test_kernel_matches_golden():
    import model_harmonics_b200 as mh
    from oracle import hypsometric_oracle as ho
    gold = np.load(GOLDEN)
    Z, D = mh.geopotential_levels(*ho.synthetic_model_levels(nlev=11, nlat=10, nlon=18, seed=20))
    assert np.array_equal(D, gold['DIFF']) and _ulp_err(Z, gold['Z']) <= 1.0


def test_model_levels_to_stokes_chain_stays_on_device():
    """T, q, sp -> (GPH, dp) -> Clm/Slm without leaving the GPU, against the oracle chain."""
    import torch
    import model_harmonics_b200 as mh
    from model_harmonics_b200 import testing as syn
    from oracle import hypsometric_oracle as ho, stokes_oracle
    nlev, nlat, nlon, L = 20, 46, 90, 40
    T, Q, SP, zs, A, B, AI, BI = ho.synthetic_model_levels(nlev, nlat, nlon, seed=9)
    lon, lat = syn.grid_axes(nlat, nlon)
    geoid = 30.0 * np.random.default_rng(1).standard_normal((nlat, nlon))
    kw = dict(LMAX=L, ELLIPSOID='WGS84', GEOID=geoid, LOVE=syn.love_numbers(L))
    # oracle chain, exactly as the drivers do it: float32 Z / GRAVITY (reanalysis_atmospheric_harmonics.py:357)
    Zo, Do = ho.geopotential_levels(T, Q, SP, zs, A, B, AI, BI)
    want = stokes_oracle.gen_atmosphere_stokes(Zo / 9.80665, Do, lon, lat, **kw)
    Zd, Dd = mh.geopotential_levels(T, Q, SP, zs, A, B, AI, BI, divide_by=9.80665, as_numpy=False)
    assert Zd.is_cuda and Zd.dtype == torch.float32
    got = mh.gen_atmosphere_stokes(Zd, Dd, lon, lat, **kw)
    assert rel_to_max(got.clm, got.slm, want.clm, want.slm) <= TOL
