"""
NumPy restatement of the drivers' per-step pre/post operations (TEST INFRASTRUCTURE).
Each function follows the cited lines of the reference drivers; they need netCDF files to run as
shipped, so these few lines are restated (parity unpinned against an executed reference, but each
is a one-line NumPy expression).
"""
import numpy as np


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
