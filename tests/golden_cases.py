"""
Seeded input builders for the golden fixtures and the parity tests.  ``build(name)`` returns
``(function_name, args, kwargs)`` exactly as they are handed to the reference / oracle /
CUDA entry points.
"""
import hashlib

import numpy as np

from model_harmonics_b200 import testing as syn


def _press(nlat, nlon, L, seed=1234, lon0=0.0, **kw):
    lon, lat = syn.grid_axes(nlat, nlon, lon0=lon0)
    G, R = syn.surface_geometry(lon, lat, seed=seed)
    P = syn.pressure_fields(nlat, nlon, 1, seed=seed)[0]
    kwargs = dict(LMAX=L, LOVE=syn.love_numbers(L))
    kwargs.update(kw)
    return P, G, R, lon, lat, kwargs


def case_pressure_c1():
    P, G, R, lon, lat, kw = _press(181, 360, 60)
    return 'gen_pressure_stokes', (P, G, R, lon, lat), kw


def case_pressure_options():
    # lon x lat orientation, negative longitudes, MMAX < LMAX, explicit WEIGHT
    P, G, R, lon, lat, kw = _press(61, 120, 40, seed=7, lon0=-180.0, MMAX=20)
    th = np.radians(90.0 - lat)
    kw['WEIGHT'] = np.sin(th) * np.radians(3.0) * np.radians(3.0) * 1.25
    return 'gen_pressure_stokes', (P.T.copy(), G.T.copy(), R.T.copy(), lon, lat), kw


def case_pressure_scalar_geometry():
    # scalar gravity and 1-element radius, as in the reference's own test
    P, G, R, lon, lat, kw = _press(90, 180, 60, seed=11)
    lat = 90.0 - 1.0 - np.arange(90) * 2.0     # cell-centred, no poles
    return 'gen_pressure_stokes', (P, 9.80665, np.atleast_1d(6371000.79), lon, lat), kw


def case_pressure_highdeg():
    P, G, R, lon, lat, kw = _press(121, 240, 200, seed=5)
    return 'gen_pressure_stokes', (P, G, R, lon, lat), kw


def case_pressure_mean_field():
    # field with a dominant mean (ill-conditioned small coefficients)
    P, G, R, lon, lat, kw = _press(91, 180, 60, seed=3)
    return 'gen_pressure_stokes', (P + 1.0e5, G, R, lon, lat), kw


def _atm(nlev, nlat, nlon, L, seed=1234, **kw):
    lon, lat = syn.grid_axes(nlat, nlon)
    GPH, dP, geoid = syn.atmosphere_cube(nlev, nlat, nlon, seed=seed)
    kwargs = dict(LMAX=L, ELLIPSOID='WGS84', GEOID=geoid, LOVE=syn.love_numbers(L))
    kwargs.update(kw)
    return 'gen_atmosphere_stokes', (GPH, dP, lon, lat), kwargs


def case_atmosphere_bc05():
    return _atm(9, 91, 180, 60, METHOD='BC05')


def case_atmosphere_sw02():
    return _atm(9, 91, 180, 60, METHOD='SW02')


def case_atmosphere_levels37():
    return _atm(37, 46, 90, 40, seed=21, MMAX=12, ELLIPSOID='GRS80')


def case_atmosphere_weight():
    f, a, kw = _atm(4, 61, 120, 30, seed=8)
    th = np.radians(90.0 - a[3])
    kw['WEIGHT'] = np.sin(th) * np.radians(3.0)**2
    return f, a, kw


def _pts(n, L, seed=1234, **kw):
    P, G, R, lon, lat, AREA = syn.point_cloud(n, seed=seed)
    kwargs = dict(AREA=AREA, LMAX=L, LOVE=syn.love_numbers(L))
    kwargs.update(kw)
    return 'gen_point_pressure', (P, G, R, lon, lat), kwargs


def case_points_3000():
    return _pts(3000, 60)


def case_points_mmax0():
    return _pts(777, 30, seed=2, MMAX=0)


def case_points_mmax_trunc():
    f, a, kw = _pts(1200, 45, seed=4, MMAX=10)
    # N-D inputs are flattened; 1-element radius
    a = tuple(np.reshape(v, (30, 40)) for v in a[:2]) + (np.atleast_1d(6371000.79),) + \
        tuple(np.reshape(v, (30, 40)) for v in a[3:])
    kw['AREA'] = np.reshape(kw['AREA'], (30, 40))
    return f, a, kw


CASES = {
    'pressure_c1': case_pressure_c1,
    'pressure_options': case_pressure_options,
    'pressure_scalar_geometry': case_pressure_scalar_geometry,
    'pressure_highdeg': case_pressure_highdeg,
    'pressure_mean_field': case_pressure_mean_field,
    'atmosphere_bc05': case_atmosphere_bc05,
    'atmosphere_sw02': case_atmosphere_sw02,
    'atmosphere_levels37': case_atmosphere_levels37,
    'atmosphere_weight': case_atmosphere_weight,
    'points_3000': case_points_3000,
    'points_mmax0': case_points_mmax0,
    'points_mmax_trunc': case_points_mmax_trunc,
}


def build(name):
    return CASES[name]()


def digest(args, kwargs):
    h = hashlib.sha256()
    for v in list(args) + [kwargs[k] for k in sorted(kwargs)]:
        if isinstance(v, (tuple, list)):
            for w in v:
                h.update(np.ascontiguousarray(w).tobytes())
        elif isinstance(v, str):
            h.update(v.encode())
        else:
            h.update(np.ascontiguousarray(v).tobytes())
    return h.hexdigest()
