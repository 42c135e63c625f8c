"""
Stand-in for the four ``gravity_toolkit`` symbols the hot path uses
(TEST INFRASTRUCTURE -- see ``oracle/__init__.py``).

``gravity_toolkit`` (pinned 1.2.7, reference ``pixi.lock:183,6864-6866``) is NOT in
``/root/reference`` and cannot be installed offline.  The routines below restate the
published algorithms it implements:

* ``plm_holmes``  -- Holmes & Featherstone (2002, J. Geodesy 76:279) column recursion for
  the fully-normalised (4-pi, no Condon-Shortley phase) associated Legendre functions,
  with the 1e-280 global scale factor.  Called at reference
  ``gen_pressure_stokes.py:182`` and ``gen_atmosphere_stokes.py:198``.
* ``legendre``    -- single-degree associated Legendre functions by downward order
  recursion from the sectoral term (the algorithm of MATLAB's ``legendre``).  Called at
  ``gen_point_pressure.py:184``.
* ``units``       -- degree-dependent unit factors; only ``rad_e``, ``l`` and
  ``spatial(hl,kl,ll).mmwe`` are used (``gen_pressure_stokes.py:155-160``,
  ``gen_atmosphere_stokes.py:175-176``, ``gen_point_pressure.py:116-122``).
* ``harmonics``   -- result container (``gen_pressure_stokes.py:191-193``).

Validation (tests/test_oracle_legendre.py): plm_holmes vs scipy ``lpmv`` with the
geodesy normalisation, Gauss-Legendre orthonormality, legendre vs plm_holmes including
|lat| up to 89.99 deg and the poles.
"""
import copy as _copy
import numpy as np

__all__ = ['plm_holmes', 'legendre', 'units', 'harmonics']

_SCALEF = 1.0e-280


def _hf_coefficients(lmax, dtype):
    """Holmes-Featherstone recursion multipliers, packed by k = l(l+1)/2 + m."""
    nk = (lmax + 1) * (lmax + 2) // 2
    f1 = np.zeros(nk, dtype=dtype)
    f2 = np.zeros(nk, dtype=dtype)
    for l in range(2, lmax + 1):
        base = l * (l + 1) // 2
        tl1 = np.sqrt(2.0 * l + 1.0)
        # zonal column
        f1[base] = np.sqrt(2.0 * l - 1.0) * tl1 / dtype(l)
        f2[base] = dtype(l - 1.0) * tl1 / (np.sqrt(2.0 * l - 3.0) * dtype(l))
        for m in range(1, l - 1):
            f1[base + m] = tl1 * np.sqrt(2.0 * l - 1.0) / (np.sqrt(l + m) * np.sqrt(l - m))
            f2[base + m] = tl1 * np.sqrt(l - m - 1.0) * np.sqrt(l + m - 1.0) / \
                (np.sqrt(2.0 * l - 3.0) * np.sqrt(l + m) * np.sqrt(l - m))
    return f1, f2


def plm_holmes(LMAX, x, ASTYPE=np.float64):
    """
    Fully-normalised associated Legendre functions P_lm(x), 0 <= m <= l <= LMAX, and their
    first derivative with respect to colatitude.  Returns ``(plm, dplm)`` with shape
    ``(LMAX+1, LMAX+1, len(x))``; entries with m > l are zero.
    """
    x = np.atleast_1d(x).flatten().astype(ASTYPE)
    nx = len(x)
    LMAX = np.int64(LMAX)
    L = int(LMAX)
    idx = lambda l, m: l * (l + 1) // 2 + m
    f1, f2 = _hf_coefficients(L, ASTYPE)
    p = np.zeros(((L + 1) * (L + 2) // 2, nx), dtype=ASTYPE)
    # sin(colatitude); exact zeros are replaced by machine epsilon
    u = np.sqrt(1.0 - x**2)
    u[u == 0] = np.finfo(u.dtype).eps
    # zonal terms are computed unscaled
    p[0, :] = 1.0
    if L >= 1:
        p[1, :] = np.sqrt(3.0) * x
    for l in range(2, L + 1):
        k = idx(l, 0)
        p[k, :] = f1[k] * x * p[idx(l - 1, 0), :] - f2[k] * p[idx(l - 2, 0), :]
    # non-zonal columns: recursion on P_lm/u^m scaled by 1e-280, rescaled by u^m/1e-280 at the end
    pmm = np.sqrt(2.0) * _SCALEF
    rescalem = 1.0 / _SCALEF
    for m in range(1, L):
        rescalem = rescalem * u
        pmm = pmm * np.sqrt(2 * m + 1) / np.sqrt(2 * m)
        p[idx(m, m), :] = pmm
        p[idx(m + 1, m), :] = x * np.sqrt(2 * m + 3) * pmm
        for l in range(m + 2, L + 1):
            k = idx(l, m)
            p[k, :] = x * f1[k] * p[idx(l - 1, m), :] - f2[k] * p[idx(l - 2, m), :]
            p[idx(l - 2, m), :] = p[idx(l - 2, m), :] * rescalem
        p[idx(L, m), :] = p[idx(L, m), :] * rescalem
        p[idx(L - 1, m), :] = p[idx(L - 1, m), :] * rescalem
    if L >= 1:
        # last sectoral term
        rescalem = rescalem * u
        p[idx(L, L), :] = pmm * np.sqrt(2 * L + 1) / np.sqrt(2 * L) * rescalem
    plm = np.zeros((L + 1, L + 1, nx), dtype=ASTYPE)
    dplm = np.zeros((L + 1, L + 1, nx), dtype=ASTYPE)
    for m in range(L + 1):
        for l in range(m, L + 1):
            plm[l, m, :] = p[idx(l, m), :]
            if l == m:
                dplm[l, m, :] = ASTYPE(m) * (x / u) * plm[l, m, :]
            else:
                flm = np.sqrt(((l**2.0 - m**2.0) * (2.0 * l + 1.0)) / (2.0 * l - 1.0))
                dplm[l, m, :] = (1.0 / u) * (l * x * plm[l, m, :] - flm * plm[l - 1, m, :])
    return plm, dplm


def legendre(l, x, NORMALIZE=False):
    """
    Associated Legendre functions of a single degree ``l`` for orders 0..l, shape
    ``(l+1, len(x))``.  ``NORMALIZE=True`` gives the same 4-pi normalised, phase-free
    functions as ``plm_holmes``.
    """
    x = np.atleast_1d(x).flatten()
    nx = len(x)
    if l == 0:
        return np.ones((1, nx), dtype=np.float64)
    rootl = np.sqrt(np.arange(0, 2 * l + 1))
    s = np.sqrt(1.0 - x**2)
    P = np.zeros((l + 3, nx), dtype=np.float64)
    twocot = np.zeros(nx)
    nz = s > 0
    twocot[nz] = -2.0 * x[nz] / s[nz]
    sn = (-s)**l
    tol = np.sqrt(np.finfo(np.float64).tiny)
    # columns where s**l underflows: start from the first non-negligible order
    under, = np.nonzero((s > 0) & (np.abs(sn) <= tol))
    if len(under) > 0:
        v = 9.2 - np.log(tol) / (l * s[under])
        w = 1.0 / np.log(v)
        m1 = 1 + l * s[under] * v * w * (1.0058 + w * (3.819 - w * 12.173))
        m1 = np.where(l < np.floor(m1), l, np.floor(m1)).astype(np.int64)
        for j, mm1 in enumerate(m1):
            c = under[j]
            P[mm1 - 1:l + 1, c] = 0.0
            tstart = np.finfo(np.float64).eps
            P[mm1 - 1, c] = np.sign(np.fmod(mm1, 2) - 0.5) * tstart
            if x[c] < 0:
                P[mm1 - 1, c] = np.sign(np.fmod(l + 1, 2) - 0.5) * tstart
            sumsq = tol
            for m in range(mm1 - 2, -1, -1):
                P[m, c] = ((m + 1) * twocot[c] * P[m + 1, c] -
                           rootl[l + m + 2] * rootl[l - m - 1] * P[m + 2, c]) / \
                    (rootl[l + m + 1] * rootl[l - m])
                sumsq += P[m, c]**2
            scale = 1.0 / np.sqrt(2.0 * sumsq - P[0, c]**2)
            P[:mm1 + 1, c] *= scale
    # regular columns: downward recursion from the sectoral term
    reg, = np.nonzero((x != 1) & (np.abs(sn) >= tol))
    if len(reg) > 0:
        d = np.arange(2, 2 * l + 2, 2)
        c = np.prod(1.0 - 1.0 / d)
        P[l, reg] = np.sqrt(c) * sn[reg]
        P[l - 1, reg] = P[l, reg] * twocot[reg] * l / rootl[2 * l]
        for m in range(l - 2, -1, -1):
            P[m, reg] = (P[m + 1, reg] * twocot[reg] * (m + 1.0) -
                         P[m + 2, reg] * rootl[l + m + 2] * rootl[l - m - 1]) / \
                (rootl[l + m + 1] * rootl[l - m])
    Pl = np.copy(P[:l + 1, :])
    # poles
    s0, = np.nonzero(s == 0)
    if len(s0) > 0:
        Pl[0, s0] = x[s0]**l
    if NORMALIZE:
        norm = np.zeros((l + 1))
        norm[0] = np.sqrt(2.0 * l + 1)
        m = np.arange(1, l + 1)
        norm[1:] = (-1)**m * np.sqrt(2.0 * (2.0 * l + 1.0))
        Pl *= norm[:, np.newaxis]
    else:
        for m in range(1, l):
            Pl[m, :] *= np.prod(rootl[l - m + 1:l + m + 1])
        Pl[l, :] *= np.prod(rootl[1:])
    return Pl


class units(object):
    """Degree-dependent unit conversion factors (CGS), only what the hot path reads."""
    # mean density of the Earth [g/cm^3] -- unverifiable here; it is a uniform
    # scale on every coefficient and cancels in GPU-vs-oracle parity
    rho_e = 5.517

    def __init__(self, lmax=None, a_axis=6.378137e8, flat=1.0 / 298.257223563):
        self.a_axis = a_axis
        self.flat = flat
        # radius of the sphere with the volume of the ellipsoid [cm]
        self.rad_e = a_axis * (1.0 - flat)**(1.0 / 3.0)
        self.lmax = lmax
        self.l = np.arange(lmax + 1) if lmax is not None else None
        self.norm = None
        self.cmwe = None
        self.mmwe = None

    def spatial(self, hl, kl, ll, **kwargs):
        """Factors converting equivalent water thickness to normalised Stokes coefficients."""
        l = self.l
        self.norm = np.ones((self.lmax + 1))
        self.cmwe = 3.0 * (1.0 + kl[l]) / (1.0 + 2.0 * l) / (4.0 * np.pi * self.rad_e * self.rho_e)
        self.mmwe = 3.0 * (1.0 + kl[l]) / (1.0 + 2.0 * l) / (40.0 * np.pi * self.rad_e * self.rho_e)
        return self


class harmonics(object):
    """Minimal spherical-harmonic container with the methods used around the hot path."""
    np.seterr(invalid='ignore')

    def __init__(self, **kwargs):
        self.clm = None
        self.slm = None
        self.time = None
        self.month = None
        self.lmax = kwargs.get('lmax', None)
        self.mmax = kwargs.get('mmax', None)
        self.attributes = dict()
        self.filename = None
        self.update_dimensions()

    def update_dimensions(self):
        self.l = np.arange(self.lmax + 1) if self.lmax is not None else None
        self.m = np.arange(self.mmax + 1) if self.mmax is not None else None
        return self

    @property
    def shape(self):
        return np.shape(self.clm)

    @property
    def ndim(self):
        return np.ndim(self.clm)

    @property
    def amplitude(self):
        """Degree amplitude sqrt(sum_m clm^2 + slm^2)."""
        return np.sqrt(np.sum(self.clm**2 + self.slm**2, axis=1))

    def copy(self):
        return _copy.deepcopy(self)

    def zeros_like(self):
        out = harmonics(lmax=self.lmax, mmax=self.mmax)
        out.clm = np.zeros_like(self.clm)
        out.slm = np.zeros_like(self.slm)
        return out

    def _binary(self, other, op):
        l1 = min(int(self.lmax), int(other.lmax)) + 1
        m1 = min(int(self.mmax), int(other.mmax)) + 1
        self.clm[:l1, :m1] = op(self.clm[:l1, :m1], other.clm[:l1, :m1])
        self.slm[:l1, :m1] = op(self.slm[:l1, :m1], other.slm[:l1, :m1])
        return self

    def add(self, other):
        return self._binary(other, np.add)

    def subtract(self, other):
        return self._binary(other, np.subtract)

    def scale(self, var):
        out = self.copy()
        out.clm = var * self.clm
        out.slm = var * self.slm
        return out

    def truncate(self, lmax, lmin=0, mmax=None):
        mmax = lmax if mmax is None else mmax
        out = harmonics(lmax=np.int64(lmax), mmax=np.int64(mmax))
        out.clm = np.zeros((lmax + 1, mmax + 1))
        out.slm = np.zeros((lmax + 1, mmax + 1))
        l1 = min(int(self.lmax), lmax) + 1
        m1 = min(int(self.mmax), mmax) + 1
        out.clm[lmin:l1, :m1] = self.clm[lmin:l1, :m1]
        out.slm[lmin:l1, :m1] = self.slm[lmin:l1, :m1]
        out.time = _copy.copy(self.time)
        out.month = _copy.copy(self.month)
        return out

    def from_list(self, objs):
        self.lmax = objs[0].lmax
        self.mmax = objs[0].mmax
        self.update_dimensions()
        self.clm = np.stack([o.clm for o in objs], axis=-1)
        self.slm = np.stack([o.slm for o in objs], axis=-1)
        self.time = np.array([o.time for o in objs])
        self.month = np.array([o.month for o in objs])
        return self

    def mean(self):
        out = harmonics(lmax=self.lmax, mmax=self.mmax)
        out.clm = np.mean(self.clm, axis=-1)
        out.slm = np.mean(self.slm, axis=-1)
        return out
