"""
NumPy restatement of the reference's fields -> Stokes hot path (TEST INFRASTRUCTURE --
see ``oracle/__init__.py``; the product package never imports this file).

Each function follows, step for step, the algorithm of the reference file cited in its
docstring (paths relative to ``/root/reference``), with the same NumPy primitives
(per-degree ``np.power`` pass, complex ``np.einsum`` Fourier sum, per-degree Legendre
contraction), so that (i) results agree with the reference files to rounding and
(ii) its wall-clock is an honest "port" baseline of the reference's CPU cost.
It is pinned against the unmodified reference files by ``tests/test_oracle_vs_reference.py``
(build container) and against ``tests/golden/*.npz`` (anywhere).
"""
import numpy as np

from oracle.gravtk_standin import plm_holmes, legendre, units, harmonics

# ellipsoid constants read by the atmosphere path (reference model_harmonics/datum.py:121-189)
_ELLIPSOIDS = {
    'CLK66': (6378206.4, 1.0 / 294.9786982), 'NAD27': (6378206.4, 1.0 / 294.9786982),
    'GRS80': (6378135.0, 1.0 / 298.26), 'NAD83': (6378135.0, 1.0 / 298.26),
    'GRS67': (6378160.0, 1.0 / 298.247167427),
    'WGS72': (6378135.0, 1.0 / 298.26),
    'WGS84': (6378137.0, 1.0 / 298.257223563),
    'ATS77': (6378135.0, 1.0 / 298.257),
    'KRASS': (6378245.0, 1.0 / 298.3),
    'INTER': (6378388.0, 1 / 297.0),
    'MAIRY': (6377340.189, 1 / 299.3249646),
    'HGH80': (6378273.0, 1.0 / 298.279411123064),
    'TOPEX': (6378136.3, 1.0 / 298.257),
    'EGM96': (6378136.3, 1.0 / 298.256415099),
    'IERS': (6378136.6, 1.0 / 298.25642),
}


def ellipsoid_constants(name):
    """a_axis, flat, rad_e, ecc1 as in reference datum.py:114-219,241-249."""
    a_axis, flat = _ELLIPSOIDS[name.upper()]
    rad_e = a_axis * (1.0 - flat)**(1.0 / 3.0)
    ecc = np.sqrt((2.0 * flat - flat**2) * a_axis**2)
    return a_axis, flat, rad_e, ecc / a_axis


def _weights(phi, th, WEIGHT):
    # reference gen_pressure_stokes.py:163-172 / gen_atmosphere_stokes.py:179-188
    w = np.zeros((len(th)))
    if WEIGHT is not None:
        w[:] = np.broadcast_to(np.atleast_1d(WEIGHT), len(th))
    else:
        w[:] = np.sin(th) * np.abs(phi[1] - phi[0]) * np.abs(th[1] - th[0])
    return w


def _euler_table(mmax, phi):
    # reference gen_pressure_stokes.py:176-177
    orders = np.arange(mmax + 1)
    return np.exp(1j * np.einsum('m...,p...->mp...', orders, phi))


def _contract(out, l, mmax, eul, field_hp_axes, plm_w, scale):
    # reference gen_pressure_stokes.py:201-206 (Fourier sum, Legendre sum, scaling)
    top = int(np.min([mmax, l]))
    d = np.einsum(field_hp_axes, eul, scale[0])
    y = np.einsum('mh...,mh...->m...', plm_w[l, :top + 1, :], d[:top + 1, :])
    out.clm[l, :top + 1] = scale[1] * y.real
    out.slm[l, :top + 1] = scale[1] * y.imag


def gen_pressure_stokes(P, G, R, lon, lat, LMAX=60, MMAX=None, WEIGHT=None, PLM=None, LOVE=None):
    """Follows reference model_harmonics/gen_pressure_stokes.py:81-209."""
    LMAX = np.int64(LMAX)
    MMAX = np.copy(LMAX) if not MMAX else MMAX                       # :133-135
    phi = np.radians(np.squeeze(lon))                                # :138
    th = np.radians(90.0 - np.squeeze(lat))                          # :139
    phi = np.where(phi < 0, phi + 2.0 * np.pi, phi)                  # :141
    nlat = len(th)
    if np.shape(P)[0] == nlat:                                       # :146-151
        P, G, R = np.transpose(P), np.transpose(G), np.transpose(R)
    fac = units(lmax=LMAX).spatial(*LOVE)                            # :155
    rad_e = fac.rad_e / 100.0                                        # :158
    dfactor = fac.mmwe                                               # :160
    w = _weights(phi, th, WEIGHT)
    eul = _euler_table(MMAX, phi)
    if PLM is None:
        PLM, _ = plm_holmes(LMAX, np.cos(th))                        # :180-182
    plm_w = np.einsum('lmh...,h...->lmh...', PLM[:LMAX + 1, :MMAX + 1, :], w)   # :186-188
    out = harmonics(lmax=LMAX, mmax=MMAX)
    out.clm = np.zeros((LMAX + 1, MMAX + 1))
    out.slm = np.zeros((LMAX + 1, MMAX + 1))
    for l in range(0, LMAX + 1):                                     # :194-206
        pf = (P / G) * np.power(R / rad_e, (l + 2))
        _contract(out, l, MMAX, eul, 'mp...,ph...->mh...', plm_w, (pf, dfactor[l]))
    return out


def gen_atmosphere_stokes(GPH, pressure, lon, lat, LMAX=60, MMAX=None, ELLIPSOID=None,
                          GEOID=None, WEIGHT=None, PLM=None, LOVE=None, METHOD='BC05'):
    """Follows reference model_harmonics/gen_atmosphere_stokes.py:86-278."""
    LMAX = np.int64(LMAX)
    MMAX = np.copy(LMAX) if not MMAX else MMAX                       # :147-149
    nlev, nlat, nlon = np.shape(GPH)                                 # :152
    phi = np.radians(np.squeeze(lon))
    th = np.radians(90.0 - np.squeeze(lat))
    gphi, gth = np.meshgrid(phi, th)                                 # :157
    a_axis, flat, rad_e, ecc1 = ellipsoid_constants(ELLIPSOID)       # :160-168
    N = a_axis / np.sqrt(1.0 - ecc1**2.0 * np.cos(gth)**2.0)         # :171
    dfactor = units(lmax=LMAX, a_axis=100.0 * a_axis, flat=flat).spatial(*LOVE).mmwe   # :175-176
    w = _weights(phi, th, WEIGHT)
    eul = _euler_table(MMAX, phi)
    if PLM is None:
        PLM, _ = plm_holmes(LMAX, np.cos(th))
    plm_w = np.einsum('lmh...,h...->lmh...', PLM[:LMAX + 1, :MMAX + 1, :], w)
    g0 = 9.80665
    ge = 9.780356
    gs = ge * (1.0 + 5.2885e-3 * np.cos(gth)**2 - 5.9e-6 * np.cos(2.0 * gth)**2)   # :211-215
    Rlev = np.zeros((nlev, nlat, nlon))
    gam = np.zeros((nlev, nlat, nlon))
    for k in range(nlev):                                            # :219-238
        H = (1.0 - 0.002644 * np.cos(2.0 * gth)) * GPH[k, :, :] + \
            (1.0 - 0.0089 * np.cos(2.0 * gth)) * (GPH[k, :, :]**2) / 6.245e6
        X = (N + GEOID + H) * np.sin(gth) * np.cos(gphi)
        Y = (N + GEOID + H) * np.sin(gth) * np.sin(gphi)
        Z = (N * (1.0 - ecc1**2.0) + GEOID + H) * np.cos(gth)
        Rlev[k, :, :] = np.sqrt(X**2.0 + Y**2.0 + Z**2.0)
        gam[k, :, :] = gs * (1.0 - 2.0 * (1.006803 - 0.06706 * np.cos(gth)**2) * (H / rad_e) +
                             3.0 * (H / rad_e)**2)
    pf = np.empty((nlat, nlon))
    out = harmonics(lmax=LMAX, mmax=MMAX)
    out.clm = np.zeros((LMAX + 1, MMAX + 1))
    out.slm = np.zeros((LMAX + 1, MMAX + 1))
    for l in range(0, LMAX + 1):                                     # :246-275
        pf[:, :] = 0.0
        for k in range(nlev):
            if METHOD == 'SW02':                                     # :254-260
                pf += (pressure[k, :, :] / g0) * np.power(
                    rad_e / (rad_e - GPH[k, :, :]) + (GEOID / rad_e), (l + 4))
            elif METHOD == 'BC05':                                   # :261-265
                pf += (pressure[k, :, :] / gam[k, :, :]) * np.power(Rlev[k, :, :] / rad_e, (l + 2))
        _contract(out, l, MMAX, eul, 'mp...,hp...->mh...', plm_w, (-pf, dfactor[l]))   # :270-275
    return out


def gen_point_pressure(P, G, R, lon, lat, AREA=None, LMAX=60, MMAX=None, LOVE=None):
    """Follows reference model_harmonics/gen_point_pressure.py:64-194."""
    if MMAX is None:                                                 # :107-108
        MMAX = np.copy(LMAX)
    phi = np.radians(lon.flatten())                                  # :111-113
    theta = np.radians(90.0 - lat.flatten())
    npts = len(phi)
    fac = units(lmax=LMAX).spatial(*LOVE)                            # :116
    rad_e = fac.rad_e / 100.0
    dfactor = 4.0 * np.pi * fac.mmwe / (1.0 + 2.0 * fac.l)           # :122
    alpha = 1.0 - AREA.flatten() / (2.0 * np.pi * rad_e**2)          # :126
    Pm1 = np.ones((npts))
    Pl = np.ones((npts))
    out = harmonics(lmax=LMAX, mmax=MMAX)
    out.clm = np.zeros((LMAX + 1, MMAX + 1))
    out.slm = np.zeros((LMAX + 1, MMAX + 1))
    x = np.cos(theta)
    for l in range(LMAX + 1):                                        # :136-154
        m1 = int(np.min([l, MMAX])) + 1
        Pp1 = ((2.0 * l + 1.0) / (l + 1.0)) * alpha * Pl - (l / (l + 1.0)) * Pm1   # :139
        Pdisc = (Pm1 - Pp1) / 2.0                                    # :142
        PGR = Pdisc * (P.flatten() / G.flatten()) * np.power(R.flatten() / rad_e, (l + 2))   # :144-147
        # degree-l harmonics of the point data (:160-194)
        leg = legendre(l, x, NORMALIZE=True)
        eul = np.exp(1j * np.einsum('m...,p...->mp...', np.arange(0, l + 1), phi))
        Yl = dfactor[l] * np.einsum('mp...,mp...,mp...->m...',
                                    np.kron(np.ones((l + 1, 1)), PGR[np.newaxis, :]), leg, eul)
        out.clm[l, :m1] = Yl.real[:m1]
        out.slm[l, :m1] = Yl.imag[:m1]
        Pm1[:] = np.copy(Pl)                                         # :153-154
        Pl[:] = np.copy(Pp1)
    return out


def gen_stokes(data, lon, lat, LMIN=0, LMAX=60, MMAX=None, UNITS=1, PLM=None, LOVE=None, WEIGHT=None):
    """
    Restatement of the third-party ``gravity_toolkit.gen_stokes`` (pinned 1.2.7, not vendored under
    /root/reference; call sites ``reanalysis/reanalysis_monthly_harmonics.py:434``,
    ``TWS/gldas_monthly_harmonics.py:373``): cos/sin(m phi) dot products over longitude, then the
    Legendre sum with ``plm * int_fact`` per degree and the degree-dependent unit factor.  PARITY
    UNPINNED against the real package; pinned instead to ``gen_pressure_stokes`` above (itself pinned
    to the reference), which it must equal for R = rad_e, G = 1, UNITS = 3
    (``tests/test_oracle_vs_reference.py``).  ``WEIGHT`` is the keyword the reanalysis driver passes.
    """
    LMAX = np.int64(LMAX)
    MMAX = np.copy(LMAX) if (MMAX is None) else MMAX
    lon, lat = np.squeeze(lon), np.squeeze(lat)
    nlat = len(lat)
    phi = lon * np.pi / 180.0
    th = (90.0 - lat) * np.pi / 180.0
    if np.shape(data)[0] == nlat:
        data = np.transpose(data)                                    # (lon, lat)
    m = np.arange(MMAX + 1)
    ccos = np.cos(np.dot(m[:, np.newaxis], phi[np.newaxis, :]))
    ssin = np.sin(np.dot(m[:, np.newaxis], phi[np.newaxis, :]))
    int_fact = _weights(phi, th, WEIGHT)
    fac = units(lmax=LMAX).spatial(*LOVE)
    if isinstance(UNITS, (list, tuple, np.ndarray)):
        dfactor = np.array(UNITS, dtype=np.float64)
    elif UNITS == 1:
        dfactor = fac.cmwe
    elif UNITS == 2:
        dfactor = fac.cmwe
        int_fact = np.full(nlat, 1.0e15 / (fac.rad_e**2))
    elif UNITS == 3:
        dfactor = fac.mmwe
    else:
        raise ValueError('Unknown units %r' % (UNITS,))
    if PLM is None:
        PLM, _ = plm_holmes(LMAX, np.cos(th))
    plm = np.zeros((LMAX + 1, MMAX + 1, nlat))
    for j in range(nlat):
        plm[:, :, j] = PLM[:LMAX + 1, :MMAX + 1, j] * int_fact[j]
    dcos = np.dot(ccos, data)
    dsin = np.dot(ssin, data)
    out = harmonics(lmax=LMAX, mmax=MMAX)
    out.clm = np.zeros((LMAX + 1, MMAX + 1))
    out.slm = np.zeros((LMAX + 1, MMAX + 1))
    for l in range(int(LMIN), LMAX + 1):
        mm = int(np.min([MMAX, l]))
        out.clm[l, :mm + 1] = dfactor[l] * np.sum(plm[l, :mm + 1, :] * dcos[:mm + 1, :], axis=1)
        out.slm[l, :mm + 1] = dfactor[l] * np.sum(plm[l, :mm + 1, :] * dsin[:mm + 1, :], axis=1)
    return out
