"""
Condition-aware error scale and extended-precision truth for the grid operators (TEST
INFRASTRUCTURE; SURVEY.md 8c secondary metric).

``A_lm = |D_l| * sum_{h,p} |w_h Plm(x_h) xi_l(h,p)|`` is the sum of the absolute values of the terms
a coefficient is made of (the cos/sin factor is bounded by one): the natural rounding-error scale
of the sum, and the strictest bound any fp64 summation order -- including the reference's own --
can be expected to meet.  Gate: ``|delta_lm| <= 1e-12 * A_lm`` per coefficient.

``pressure_truth`` evaluates gen_pressure_stokes' formula in ``np.longdouble`` (80-bit on x86) so
that the literal per-coefficient relative error of the CUDA path can be printed next to the
reference's own.
"""
import numpy as np

from oracle.gravtk_standin import plm_holmes, units


def _prep(lon, lat, LMAX, MMAX, WEIGHT, LOVE, wrap=True):
    LMAX = int(LMAX)
    MMAX = LMAX if not MMAX else int(MMAX)
    phi = np.radians(np.squeeze(lon))
    th = np.radians(90.0 - np.squeeze(lat))
    if wrap:
        phi = np.where(phi < 0, phi + 2.0 * np.pi, phi)
    if WEIGHT is not None:
        w = np.broadcast_to(np.atleast_1d(WEIGHT), len(th)).astype(np.float64)
    else:
        w = np.sin(th) * np.abs(phi[1] - phi[0]) * np.abs(th[1] - th[0])
    return LMAX, MMAX, phi, th, w


def pressure_error_scale(P, G, R, lon, lat, LMAX=60, MMAX=None, WEIGHT=None, LOVE=None):
    """A_lm for gen_pressure_stokes (gen_pressure_stokes.py:194-206), shape (LMAX+1, MMAX+1)."""
    LMAX, MMAX, phi, th, w = _prep(lon, lat, LMAX, MMAX, WEIGHT, LOVE)
    nlat = len(th)
    if np.shape(P)[0] != nlat:
        P, G, R = np.transpose(P), np.transpose(G), np.transpose(R)
    fac = units(lmax=LMAX).spatial(*LOVE)
    rad_e = fac.rad_e / 100.0
    plm, _ = plm_holmes(LMAX, np.cos(th))
    aplm = np.abs(plm[:, :MMAX + 1, :]) * np.abs(w)[None, None, :]
    f = np.abs(np.broadcast_to(P / G, (nlat, len(phi))))
    r = np.broadcast_to(np.asarray(R) / rad_e, (nlat, len(phi)))
    A = np.zeros((LMAX + 1, MMAX + 1))
    for l in range(LMAX + 1):
        ring = np.sum(f * np.power(r, l + 2), axis=1)           # sum over longitudes of |xi_l|
        A[l, :] = np.abs(fac.mmwe[l]) * (aplm[l] @ ring)
    return A


def pressure_truth(P, G, R, lon, lat, LMAX=60, MMAX=None, WEIGHT=None, LOVE=None):
    """gen_pressure_stokes' formula in extended precision; returns (clm, slm) as longdouble."""
    LD = np.longdouble
    LMAX, MMAX, phi, th, w = _prep(lon, lat, LMAX, MMAX, WEIGHT, LOVE)
    nlat = len(th)
    if np.shape(P)[0] != nlat:
        P, G, R = np.transpose(P), np.transpose(G), np.transpose(R)
    fac = units(lmax=LMAX).spatial(*LOVE)
    rad_e = LD(fac.rad_e) / LD(100.0)
    # inputs are the same binary64 values; everything downstream is extended
    phi, th, w = phi.astype(LD), th.astype(LD), w.astype(LD)
    plm, _ = plm_holmes(LMAX, np.cos(th), ASTYPE=LD)
    f = np.broadcast_to(np.asarray(P, dtype=LD) / np.asarray(G, dtype=LD), (nlat, len(phi)))
    r = np.broadcast_to(np.asarray(R, dtype=LD) / rad_e, (nlat, len(phi)))
    m = np.arange(MMAX + 1).astype(LD)
    # the reference rounds m*phi to binary64 before the trig call: keep that argument
    arg = (np.arange(MMAX + 1)[:, None] * phi.astype(np.float64)[None, :]).astype(LD)
    C, S = np.cos(arg), np.sin(arg)
    clm = np.zeros((LMAX + 1, MMAX + 1), dtype=LD)
    slm = np.zeros((LMAX + 1, MMAX + 1), dtype=LD)
    for l in range(LMAX + 1):
        xi = f * np.power(r, l + 2)
        dc, ds = C @ xi.T, S @ xi.T                     # (m, lat)
        top = min(l, MMAX)
        wp = plm[l, :top + 1, :] * w[None, :]
        clm[l, :top + 1] = LD(fac.mmwe[l]) * np.sum(wp * dc[:top + 1], axis=1)
        slm[l, :top + 1] = LD(fac.mmwe[l]) * np.sum(wp * ds[:top + 1], axis=1)
    return clm, slm
