"""
Independent validation of the third-party Legendre restatement (oracle side only):
scipy lpmv with the geodesy normalisation, Gauss-Legendre orthonormality, and the
single-degree routine against the Holmes-Featherstone table up to the poles.
"""
import numpy as np
import pytest
from scipy.special import lpmv, gammaln

from oracle.gravtk_standin import plm_holmes, legendre, units


def _norm(l, m):
    k = 1.0 if m == 0 else 2.0
    return np.sqrt(k * (2.0 * l + 1.0) * np.exp(gammaln(l - m + 1) - gammaln(l + m + 1)))


def test_plm_holmes_vs_scipy():
    L = 60
    x = np.cos(np.radians(np.linspace(0.5, 179.5, 180)))
    plm, _ = plm_holmes(L, x)
    worst = 0.0
    for l in range(L + 1):
        for m in range(l + 1):
            ref = (-1.0)**m * _norm(l, m) * lpmv(m, l, x)
            scale = np.abs(ref).max()
            worst = max(worst, np.abs(plm[l, m] - ref).max() / scale)
    assert worst < 2e-13


def test_plm_holmes_orthonormal():
    L = 40
    x, w = np.polynomial.legendre.leggauss(128)
    plm, _ = plm_holmes(L, x)
    for m in (0, 1, 7, 40):
        for l in range(m, L + 1):
            val = 0.5 * np.sum(w * plm[l, m]**2)
            assert abs(val - (1.0 if m == 0 else 2.0)) < 1e-11
    # low-degree closed forms
    assert np.allclose(plm[1, 0], np.sqrt(3.0) * x)
    assert np.allclose(plm[1, 1], np.sqrt(3.0) * np.sqrt(1 - x**2))
    # upper triangle is zero
    assert np.all(plm[3, 4:] == 0)


@pytest.mark.parametrize('L', [1, 2, 30, 60, 120])
def test_legendre_matches_plm_holmes(L):
    lat = np.concatenate([np.linspace(-89.0, 89.0, 90), [89.5, 89.9, 89.99, -89.5, -89.9, -89.99]])
    x = np.cos(np.radians(90.0 - lat))
    plm, _ = plm_holmes(L, x)
    for l in (L, max(L - 1, 1)):
        leg = legendre(l, x, NORMALIZE=True)
        assert leg.shape == (l + 1, len(x))
        assert np.abs(leg - plm[l, :l + 1]).max() < 5e-12


def test_legendre_poles_and_degree_zero():
    x = np.array([1.0, -1.0, 0.3])
    assert np.all(legendre(0, x, NORMALIZE=True) == 1.0)
    leg = legendre(5, x, NORMALIZE=True)
    assert np.isclose(leg[0, 0], np.sqrt(11.0)) and np.isclose(leg[0, 1], -np.sqrt(11.0))
    assert np.all(leg[1:, :2] == 0.0)


def test_units_mmwe():
    L = 10
    kl = np.linspace(0, -0.2, L + 1)
    f = units(lmax=L).spatial(None, kl, None)
    assert np.isclose(f.rad_e, 6.371000790009159e8)
    l = np.arange(L + 1)
    assert np.allclose(f.mmwe, 3 * (1 + kl) / (1 + 2 * l) / (40 * np.pi * f.rad_e * 5.517))
    assert np.allclose(f.cmwe, 10 * f.mmwe)
