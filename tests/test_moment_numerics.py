"""
Algorithm-level check (CPU, NumPy) of the binomial-moment form the atmosphere kernel uses at
LMAX <= 63 (model_harmonics_b200/csrc/stokes_grid.cu, produce_atmosphere<MOM> and moments_to_chunk):

    sum_k q_k r_k^n  =  r0^n sum_j C(n, j) M_j,    M_j = sum_k q_k d_k^j,   d_k = (r_k - r0)/r0

with the kernel's constants (16 moments, r0 = 1.00276).  It pins the validity range the design
document states: rounding-level agreement with the direct powers for level heights up to ~100 km,
and a truncation error that stays below the 1e-12 contract there.  The kernel itself is checked
against the oracle by the GPU tests.
"""
import math

import numpy as np
import pytest

NMOM = 16
R0 = 1.00276
RAD_E = 6371000.79


def _direct(q, r, n, dtype):
    return (q.astype(dtype) * r.astype(dtype) ** n).sum(axis=0)


def _moments(q, r, n):
    d = (r - R0) * (1.0 / R0)
    t = q.copy()
    m = np.empty((NMOM,) + q.shape[1:])
    for j in range(NMOM):
        m[j] = t.sum(axis=0)
        t = t * d
    coef = np.array([float(math.comb(n, j)) * R0 ** n if j <= n else 0.0 for j in range(NMOM)])
    return np.tensordot(coef, m, axes=(0, 0))


@pytest.mark.parametrize('top_km', [50.0, 80.0, 100.0])
def test_moment_form_matches_direct_powers(top_km):
    rng = np.random.default_rng(int(top_km))
    nlev, ncol = 37, 400
    # radius ratios of an ellipsoidal Earth with topography and levels up to top_km
    base = 1.0 + rng.uniform(-21.8e3, 9.0e3, ncol) / RAD_E
    height = np.sort(rng.uniform(0.0, top_km * 1e3, (nlev, ncol)), axis=0) / RAD_E
    r = base[None, :] + height
    q = rng.uniform(0.2, 3.0, (nlev, ncol)) * 1.0e2
    for n in (2, 34, 62, 64):                       # l + 2 for l = 0, 32, 60 and l + 4 for l = 60
        truth = _direct(q, r, n, np.longdouble)
        scale = np.abs(truth).max()
        e_mom = float(np.abs(_moments(q, r, n) - truth).max() / scale)
        e_dir = float(np.abs(_direct(q, r, n, np.float64) - truth).max() / scale)
        assert e_mom <= 1.0e-13, (n, e_mom)
        assert e_mom <= 20.0 * max(e_dir, 1.0e-16), (n, e_mom, e_dir)


def test_truncation_bound_of_the_design_document():
    # (n |d|)^16 / 16!  for the ranges DESIGN.md quotes
    bound = lambda n, d: (n * d) ** NMOM / math.factorial(NMOM)
    assert bound(64, 0.007) < 1.0e-16          # levels to ~50 km
    assert bound(64, 0.016) < 1.0e-13          # levels to ~100 km
