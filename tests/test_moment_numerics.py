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


# ---------------------------------------------------------------------------------------------------
# Contracted-moment form (model_harmonics_b200/csrc/stokes_moments.cu): the binomial expansion is applied
# AFTER the Fourier and Legendre sums, around a per-ring expansion point r0(h):
#   C_lm = D_l sum_j C(l + n0, j) sum_h w_h P_lm(x_h) r0(h)^(l + n0) sum_p e^{i m phi_p} M_j[h, p]
# (BC05 takes its moments in e = (r / r0)^2 - 1 with the generalised coefficients C((l + 2) / 2, j)).
# Emulated here in NumPy, in the kernels' order of summation, against the oracle.
# ---------------------------------------------------------------------------------------------------
def _binom(nmom, L, off, half=False):
    """C(a, j), a = l + off, or the generalised C((l + off) / 2, j) of the expansion in e = (1 + d)^2 - 1."""
    E = np.zeros((nmom, L + 1))
    for l in range(L + 1):
        a = 0.5 * (l + off) if half else float(l + off)
        c = 1.0
        for j in range(nmom):
            if j > 0:
                c = c * (a - (j - 1)) / j
            E[j, l] = c
    return E


def _contract_moments(M, r0, w, PLM, eul, dfactor, L, off, scale, half=False):
    F = np.einsum('mp,jhp->hmj', eul, M)
    E = _binom(M.shape[0], L, off, half)
    clm = np.zeros((L + 1, L + 1))
    slm = np.zeros((L + 1, L + 1))
    for l in range(L + 1):
        pw = w * r0**(l + off)
        G = np.einsum('mh,hmj->mj', PLM[l, :l + 1, :] * pw[None, :], F[:, :l + 1, :])
        y = scale * dfactor[l] * (G @ E[:, l])
        clm[l, :l + 1], slm[l, :l + 1] = y.real, y.imag
    return clm, slm


def _atmosphere_emulation(GPH, dP, lon, lat, L, geoid, LOVE, method, hc=24000.0):
    from oracle import stokes_oracle as so
    from oracle.gravtk_standin import plm_holmes, units
    phi, th = np.radians(lon), np.radians(90.0 - lat)
    a_axis, flat, rad_e, ecc1 = so.ellipsoid_constants('WGS84')
    dfactor = units(lmax=L, a_axis=100.0 * a_axis, flat=flat).spatial(*LOVE).mmwe
    w = np.sin(th) * np.abs(phi[1] - phi[0]) * np.abs(th[1] - th[0])
    PLM, _ = plm_holmes(L, np.cos(th))
    st, ct = np.sin(th), np.cos(th)
    N = a_axis / np.sqrt(1.0 - ecc1**2 * ct**2)
    M = np.zeros((NMOM,) + GPH.shape[1:])
    dmax = 0.0
    if method == 'BC05':
        off, scale = 2, -1.0
        r0 = np.sqrt((N * st)**2 + (N * (1 - ecc1**2) * ct)**2) / rad_e + hc / rad_e
        a1 = 1.0 - 0.002644 * np.cos(2.0 * th)
        a2 = (1.0 - 0.0089 * np.cos(2.0 * th)) / 6.245e6
        gs = 9.780356 * (1.0 + 5.2885e-3 * ct**2 - 5.9e-6 * np.cos(2.0 * th)**2)
        g1 = -gs * 2.0 * (1.006803 - 0.06706 * ct**2) / rad_e
        g2 = 3.0 * gs / rad_e**2
        for k in range(GPH.shape[0]):
            z = GPH[k]
            H = z * (a1[:, None] + a2[:, None] * z)
            S = ((N[:, None] + geoid + H) * st[:, None])**2 + ((N[:, None] * (1 - ecc1**2) + geoid + H) * ct[:, None])**2
            q = dP[k] / (gs[:, None] + H * (g1[:, None] + g2[:, None] * H))
            d = S / (rad_e * r0[:, None])**2 - 1.0          # e = (r / r0)^2 - 1: no square root
            dmax = max(dmax, 0.5 * float(np.abs(d).max()))
            t = q.copy()
            for j in range(NMOM):
                M[j] += t
                t = t * d
    else:
        off, scale = 4, -1.0 / 9.80665
        r0 = np.full(len(th), 1.0 + hc / rad_e)
        for k in range(GPH.shape[0]):
            d = (rad_e / (rad_e - GPH[k]) + geoid / rad_e) / r0[:, None] - 1.0
            dmax = max(dmax, float(np.abs(d).max()))
            t = dP[k].copy()
            for j in range(NMOM):
                M[j] += t
                t = t * d
    eul = np.exp(1j * np.outer(np.arange(L + 1), phi))
    return _contract_moments(M, r0, w, PLM, eul, dfactor, L, off, scale, half=(method == 'BC05')) + (dmax,)


@pytest.mark.parametrize('method', ['BC05', 'SW02'])
@pytest.mark.parametrize('top_scale', [1.0, 80.0 / 48.0])
def test_contracted_moment_atmosphere_matches_oracle(method, top_scale):
    from oracle import stokes_oracle as so
    from model_harmonics_b200 import testing as syn
    L = 60
    LOVE = syn.love_numbers(L)
    lon, lat = syn.grid_axes(46, 90)
    GPH, dP, geoid = syn.atmosphere_cube(37, 46, 90, seed=5)
    GPH = GPH * top_scale
    ref = so.gen_atmosphere_stokes(GPH, dP, lon, lat, LMAX=L, ELLIPSOID='WGS84', GEOID=geoid, LOVE=LOVE, METHOD=method)
    clm, slm, dmax = _atmosphere_emulation(GPH, dP, lon, lat, L, geoid, LOVE, method)
    # inside the guard of the kernel (stokes_moments.cu: n_max |d| <= 0.8)
    assert dmax * (L + (2 if method == 'BC05' else 4)) <= 0.8
    scale = max(np.abs(ref.clm).max(), np.abs(ref.slm).max())
    err = max(np.abs(clm - ref.clm).max(), np.abs(slm - ref.slm).max()) / scale
    assert err <= 1.0e-13, err


def test_contracted_moment_pressure_lmax360_matches_oracle():
    """Per-ring expansion point = mid-range of R / rad_e over the ring (pressure_moment_kernel)."""
    from oracle import stokes_oracle as so
    from oracle.gravtk_standin import plm_holmes, units
    from model_harmonics_b200 import testing as syn
    L = 360
    LOVE = syn.love_numbers(L)
    lon, lat = syn.grid_axes(31, 60)
    G, R = syn.surface_geometry(lon, lat, seed=5)
    P = syn.pressure_fields(31, 60, 1, seed=5)[0] + 1.0e4
    ref = so.gen_pressure_stokes(P, G, R, lon, lat, LMAX=L, LOVE=LOVE)
    phi, th = np.radians(lon), np.radians(90.0 - lat)
    fac = units(lmax=L).spatial(*LOVE)
    rad_e = fac.rad_e / 100.0
    w = np.sin(th) * np.abs(phi[1] - phi[0]) * np.abs(th[1] - th[0])
    PLM, _ = plm_holmes(L, np.cos(th))
    r = R / rad_e
    r0 = 0.5 * (r.max(axis=1) + r.min(axis=1))
    d = r / r0[:, None] - 1.0
    assert np.abs(d).max() * (L + 2) <= 0.8
    M = np.zeros((NMOM,) + P.shape)
    t = P / G
    for j in range(NMOM):
        M[j] = t
        t = t * d
    eul = np.exp(1j * np.outer(np.arange(L + 1), phi))
    clm, slm = _contract_moments(M, r0, w, PLM, eul, fac.mmwe, L, 2, 1.0)
    scale = max(np.abs(ref.clm).max(), np.abs(ref.slm).max())
    err = max(np.abs(clm - ref.clm).max(), np.abs(slm - ref.slm).max()) / scale
    assert err <= 1.0e-13, err


def test_guard_bound_of_the_contracted_moment_kernels():
    # truncated tail (n |d|)^16 / 16! at the guard n |d| = 0.8
    assert 0.8 ** NMOM / math.factorial(NMOM) < 2.0e-15
