"""
Deterministic synthetic inputs for the fields -> Stokes hot path (host NumPy only).

Shared by ``tests/``, ``tests/golden/make_golden.py`` and ``bench.py`` so that the CUDA
path, the oracle and the unmodified reference all consume identical arrays.  Recipes follow
SURVEY.md section 8(d); every generator is seeded.
"""
import numpy as np

__all__ = ['love_numbers', 'grid_axes', 'surface_geometry', 'pressure_fields',
           'point_cloud', 'atmosphere_cube', 'atmosphere_field_scalars', 'model_level_inputs']

_WGS84_A = 6378137.0
_WGS84_F = 1.0 / 298.257223563


def love_numbers(lmax):
    """Smooth synthetic load Love numbers (hl, kl, ll); no data file is available offline."""
    l = np.arange(lmax + 1, dtype=np.float64)
    kl = -0.3 / (1.0 + l / 10.0)
    kl[0] = 0.0
    if lmax >= 1:
        kl[1] = 0.021
    return (-np.ones(lmax + 1), kl, np.zeros(lmax + 1))


def grid_axes(nlat, nlon, lon0=0.0):
    """Pole-to-pole north->south latitudes and 0..360 (or lon0-based) longitudes, ERA5 layout."""
    lat = np.linspace(90.0, -90.0, nlat)
    lon = lon0 + np.arange(nlon) * (360.0 / nlon)
    return lon, lat


def surface_geometry(lon, lat, seed=1234):
    """
    Orography -> geocentric radius R and normal gravity G on the (nlat, nlon) grid, formed
    the way the reanalysis pressure driver does (geodetic -> Cartesian at height h on WGS84).
    """
    rng = np.random.default_rng(seed)
    nlat, nlon = len(lat), len(lon)
    h = np.clip(2000.0 * rng.standard_normal((nlat, nlon)), -100.0, 8000.0)
    geoid = 30.0 * rng.standard_normal((nlat, nlon))
    gridlon, gridlat = np.meshgrid(np.radians(lon), np.radians(lat))
    ecc2 = 2.0 * _WGS84_F - _WGS84_F**2
    N = _WGS84_A / np.sqrt(1.0 - ecc2 * np.sin(gridlat)**2)
    X = (N + h + geoid) * np.cos(gridlat) * np.cos(gridlon)
    Y = (N + h + geoid) * np.cos(gridlat) * np.sin(gridlon)
    Z = (N * (1.0 - ecc2) + h + geoid) * np.sin(gridlat)
    R = np.sqrt(X**2 + Y**2 + Z**2)
    th = np.radians(90.0 - gridlat * 180.0 / np.pi)
    G = 9.780356 * (1.0 + 5.2885e-3 * np.cos(th)**2 - 5.9e-6 * np.cos(2.0 * th)**2) * \
        (1.0 - 2.0 * h / 6.371e6)
    return G, R


def pressure_fields(nlat, nlon, nfields, seed=1234, mean=0.0):
    """nfields surface-pressure anomaly fields, 500 Pa standard deviation, seeds seed+t."""
    out = np.empty((nfields, nlat, nlon))
    for t in range(nfields):
        out[t] = mean + 500.0 * np.random.default_rng(seed + t).standard_normal((nlat, nlon))
    return out


def point_cloud(npts, seed=1234, pole_points=True):
    """Scattered point pressures with equal disc areas (config C2)."""
    rng = np.random.default_rng(seed)
    lat = np.degrees(np.arcsin(rng.uniform(-1.0, 1.0, npts)))
    lon = rng.uniform(-180.0, 180.0, npts)
    if pole_points and npts >= 8:
        # make sure the near-pole and exactly-polar branches are exercised
        lat[:6] = [89.5, -89.9, 89.99, -89.999, 90.0, -90.0]
    P = 500.0 * rng.standard_normal(npts)
    G = 9.8 + 0.02 * rng.standard_normal(npts)
    R = 6.371e6 + 3.0e3 * rng.standard_normal(npts)
    AREA = np.full(npts, 4.0 * np.pi * 6.371e6**2 / npts)
    return P, G, R, lon, lat, AREA


def atmosphere_cube(nlev, nlat, nlon, seed=1234, dtype=np.float64):
    """Geopotential heights, pressure differences (level, lat, lon) and geoid (lat, lon)."""
    rng = np.random.default_rng(seed)
    z0 = np.clip(1500.0 * rng.standard_normal((nlat, nlon)), 0.0, 6000.0)
    GPH = np.empty((nlev, nlat, nlon), dtype=dtype)
    dP = np.empty((nlev, nlat, nlon), dtype=dtype)
    for k in range(nlev):
        GPH[k] = z0 + 48000.0 * (k / float(nlev))**1.5 + 50.0 * rng.standard_normal((nlat, nlon))
        dP[k] = -(1.0e5 / nlev) * (1.0 + 0.05 * rng.standard_normal((nlat, nlon)))
    geoid = 30.0 * rng.standard_normal((nlat, nlon))
    return GPH, dP, geoid


def atmosphere_field_scalars(t):
    """
    Field t of the multi-decade batch is derived from the base cube with IEEE-exact
    elementwise ops: GPH_t = GPH_0 * (1 + a_t), dP_t = dP_0 * (1 + b_t).
    """
    a = 1.0e-3 * np.sin(2.0 * np.pi * t / 12.0)
    b = 5.0e-3 * np.cos(2.0 * np.pi * t / 12.0) + 1.0e-5 * t
    return 1.0 + a, 1.0 + b


def model_level_inputs(nlev, nlat, nlon, seed=1234, interfaces_extra=True):
    """Seeded synthetic hybrid-level inputs (float32 T, q, sp; float64 coefficients), bottom level first."""
    rng = np.random.default_rng(seed)
    # hybrid coefficients: interface pressures p = AI + BI*sp falling from sp to ~0
    eta = np.linspace(1.0, 0.0, nlev + 1)**1.6
    BI = eta**1.5
    AI = 1.0e5 * (eta - BI) + 5.0 * (1.0 - eta)
    AI[-1] = 0.0
    BI[-1] = 0.0
    A = 0.5 * (AI[:-1] + AI[1:])
    B = 0.5 * (BI[:-1] + BI[1:])
    if not interfaces_extra:
        AI, BI = AI[:-1], BI[:-1]
    T = (288.0 - 60.0 * (1.0 - eta[:-1])[:, None, None] + 5.0 * rng.standard_normal((nlev, nlat, nlon))).astype(np.float32)
    Q = np.abs(0.012 * eta[:-1, None, None]**3 * (1.0 + 0.3 * rng.standard_normal((nlev, nlat, nlon)))).astype(np.float32)
    SP = (1.0e5 - 2.0e4 * np.abs(rng.standard_normal((nlat, nlon)))).astype(np.float32)
    zs = (9.80665 * np.clip(1500.0 * rng.standard_normal((nlat, nlon)), 0.0, 6000.0)).astype(np.float32)
    return T, Q, SP, zs, A, B, AI, BI
