"""
Pins the hypsometric oracle: against the reference's own loop lines executed from
/root/reference (build container) and against the committed golden fixture (anywhere).
"""
import os

import numpy as np
import pytest

from oracle import hypsometric_oracle as ho

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'hypsometric.npz')
CASE = dict(nlev=11, nlat=10, nlon=18, seed=20)


@pytest.mark.skipif(not ho.reference_available(), reason='reference tree not present')
@pytest.mark.parametrize('kw', [dict(nlev=12, nlat=9, nlon=16, seed=3),
                                dict(nlev=7, nlat=5, nlon=8, seed=4, interfaces_extra=False),
                                dict(nlev=37, nlat=19, nlon=36, seed=5)])
def test_oracle_is_bit_identical_to_reference_lines(kw):
    args = ho.synthetic_model_levels(**kw)
    Z1, D1 = ho.geopotential_levels(*args)
    Z2, D2 = ho.reference_loop(*args)
    assert Z1.dtype == np.float32 and D1.dtype == np.float32
    assert np.array_equal(Z1, Z2) and np.array_equal(D1, D2)


def test_oracle_matches_golden():
    gold = np.load(GOLDEN)
    Z, D = ho.geopotential_levels(*ho.synthetic_model_levels(**CASE))
    # libm log() may differ in the last bit between machines: allow one float32 ulp on Z
    assert np.array_equal(D, gold['DIFF'])
    assert np.max(np.abs(Z.astype(np.float64) - gold['Z']) / np.spacing(np.abs(gold['Z']))) <= 1.0


def test_physical_sanity():
    T, Q, SP, zs, A, B, AI, BI = ho.synthetic_model_levels(**CASE)
    Z, D = ho.geopotential_levels(T, Q, SP, zs, A, B, AI, BI)
    assert np.all(np.diff(Z, axis=0) > 0)              # geopotential rises level by level
    assert np.all(D < 0)                               # p_upper - p_lower
    # the column of pressure differences adds up to (Pmin or top interface) - surface interface pressure
    assert np.allclose(-D.sum(axis=0), AI[0] + BI[0] * SP - 0.1, rtol=1e-5)
