"""
NumPy restatement of the drivers' per-step pre/post operations (TEST INFRASTRUCTURE).
Each function follows the cited lines of the reference drivers.  The drivers are netCDF command-line
programs and cannot be called, so ``reference_pressure_step`` / ``reference_atmosphere_step`` below
EXECUTE those very source lines, read from ``/root/reference`` at run time, in a namespace of synthetic
arrays (as ``hypsometric_oracle.reference_loop`` does); ``tests/test_oracle_vs_reference.py`` holds the
restatements to them bit for bit in the build container, and ``tests/golden/driver_ops.npz`` carries
their outputs to the GPU box.
"""
import os
import textwrap

import numpy as np

from oracle.ref_loader import REFERENCE_ROOT

_PRESSURE_DRIVER = os.path.join(REFERENCE_ROOT, 'model_harmonics', 'reanalysis', 'reanalysis_pressure_harmonics.py')
_PRESSURE_LINES = (474, 490)          # P = pressure[t] - mean ... P[ii, jj] = area-weighted ocean mean
_ATMOSPHERE_DRIVER = os.path.join(REFERENCE_ROOT, 'model_harmonics', 'reanalysis',
                                  'reanalysis_atmospheric_harmonics.py')
_ATMOSPHERE_LINES = (361, 379)        # per-level redistribution of the pressure differences
_TILE_DRIVER = os.path.join(REFERENCE_ROOT, 'model_harmonics', 'OBP', 'ecco_llc_tile_harmonics.py')
_TILE_LINES = (235, 260)              # for k in range(Nt): ... obp_Ylms.add(Ylms)


def pressure_anomaly(pressure_t, mean_pressure):
    """reanalysis/reanalysis_pressure_harmonics.py:474"""
    return pressure_t - mean_pressure[:, :]


def redistribute_ocean_2d(P, AREA, ii, jj):
    """reanalysis/reanalysis_pressure_harmonics.py:476-490"""
    P = P.copy()
    TOTAL_AREA = np.sum(AREA[ii, jj])
    P[ii, jj] = np.sum(P[ii, jj] * AREA[ii, jj]) / TOTAL_AREA
    return P


def redistribute_ocean_levels(PD, AREA, ii, jj):
    """reanalysis/reanalysis_atmospheric_harmonics.py:361-379"""
    PD = PD.copy()
    TOTAL_AREA = np.sum(AREA[ii, jj])
    for p in range(PD.shape[0]):
        NEWTONS = np.sum(PD[p, ii, jj] * AREA[ii, jj])
        PD[p, ii, jj] = NEWTONS / TOTAL_AREA
    return PD


def reference_available():
    return all(os.path.isfile(f) for f in (_PRESSURE_DRIVER, _ATMOSPHERE_DRIVER, _TILE_DRIVER))


def _lines(path, span):
    src = open(path).read().splitlines()[span[0] - 1:span[1]]
    return compile(textwrap.dedent('\n'.join(src)), path, 'exec')


def grid_cell_area(lat, nlon, a_axis=6378137.0, flat=1.0 / 298.257223563):
    """The drivers' cell areas (reanalysis_pressure_harmonics.py:385-392, :484): meridional and prime-vertical
    radii of curvature times the angular steps; returns (M, N, gridtheta, dth, dphi)."""
    th = np.radians(90.0 - lat)
    gridtheta = np.repeat(th[:, None], nlon, axis=1)
    e12 = 2.0 * flat - flat**2
    N = a_axis / np.sqrt(1.0 - e12 * np.cos(gridtheta)**2.0)
    M = a_axis * (1.0 - e12) / (1.0 - e12 * np.cos(gridtheta)**2)**1.5
    return M, N, gridtheta, np.abs(th[1] - th[0]), np.radians(360.0 / nlon)


def reference_pressure_step(pressure, mean_pressure, t, lat, ii, jj, weight=None):
    """Runs reanalysis_pressure_harmonics.py:474-490 unmodified: anomaly, then ocean redistribution
    (the WEIGHT branch when ``weight`` is given).  Returns the field handed to gen_pressure_stokes."""
    nlat, nlon = mean_pressure.shape
    M, N, gridtheta, dth, dphi = grid_cell_area(lat, nlon)
    ns = dict(np=np, pressure=pressure, mean_pressure=mean_pressure, t=t, REDISTRIBUTE=True,
              WEIGHT=(weight is not None), weight=weight, nlon=nlon, M=M, N=N, gridtheta=gridtheta, dth=dth,
              dphi=dphi, ii=ii, jj=jj)
    exec(_lines(_PRESSURE_DRIVER, _PRESSURE_LINES), ns)
    return ns['P']


def reference_atmosphere_step(PD, lat, ii, jj, weight=None):
    """Runs reanalysis_atmospheric_harmonics.py:361-379 unmodified on a copy of the (level, lat, lon) cube."""
    nlevels, nlat, nlon = PD.shape
    M, N, gridtheta, dth, dphi = grid_cell_area(lat, nlon)
    ns = dict(np=np, PD=PD.copy(), REDISTRIBUTE=True, WEIGHT=(weight is not None), weight=weight, nlon=nlon,
              nlevels=nlevels, M=M, N=N, gridtheta=gridtheta, dth=dth, dphi=dphi, ii=ii, jj=jj)
    exec(_lines(_ATMOSPHERE_DRIVER, _ATMOSPHERE_LINES), ns)
    return ns['PD']


def reference_tile_accumulation(obp, gamma_h, R, area, lon, lat, LMAX, MMAX, LOVE, gen_point_pressure, harmonics):
    """
    Runs the tile loop ecco_llc_tile_harmonics.py:235-260 unmodified: ``obp`` is a masked (Nt, j, i) array,
    ``gamma_h``/``area``/``lon``/``lat`` are (Nt, j, i) and ``R`` is indexed [k, indi, indj] exactly as the
    driver does; ``gen_point_pressure`` / ``harmonics`` are the operator and container to use (the reference
    files through ``oracle.ref_loader``, or the oracle port).
    """
    import types
    Nt = obp.shape[0]
    total = harmonics(lmax=LMAX, mmax=MMAX)
    total.clm = np.zeros((LMAX + 1, MMAX + 1))
    total.slm = np.zeros((LMAX + 1, MMAX + 1))
    ns = dict(np=np, Nt=Nt, obp_data={'OBP': obp}, VARNAME='OBP', gamma_h=gamma_h, R=R,
              invariant={'area': area, 'lon': lon}, latitude_geocentric=lat, LMAX=LMAX, MMAX=MMAX, LOVE=LOVE,
              mdlhmc=types.SimpleNamespace(gen_point_pressure=gen_point_pressure), obp_Ylms=total)
    exec(_lines(_TILE_DRIVER, _TILE_LINES), ns)
    return ns['obp_Ylms']
