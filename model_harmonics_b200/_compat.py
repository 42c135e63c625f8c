"""
Host-side pieces the operators need from outside the hot path.

* ``units`` / ``harmonics``: the real ``gravity_toolkit`` classes when that package is
  importable, otherwise local equivalents exposing what the reference's hot path and its
  callers use (reference ``gen_pressure_stokes.py:155-160,191-193``).
* ``ellipsoid_constants``: a_axis, flat, rad_e, ecc1 as ``model_harmonics/datum.py:114-219,241-249``
  derives them (only these four are read at ``gen_atmosphere_stokes.py:160-168``).
* ``plm_columns``: Holmes-Featherstone columns for a handful of colatitudes, used only to
  check whether a caller-supplied ``PLM`` table is the standard one for ``lat``.
"""
import copy as _copy
import os as _os
import re as _re

import numpy as np

try:  # pragma: no cover - not installable offline
    import gravity_toolkit as _gravtk
    units = _gravtk.units
    harmonics = _gravtk.harmonics
    HAVE_GRAVITY_TOOLKIT = True
except Exception:
    _gravtk = None
    HAVE_GRAVITY_TOOLKIT = False

    class units(object):
        """Degree-dependent unit factors (CGS); ``rad_e``, ``l``, ``spatial().mmwe``."""
        # mean density of the Earth [g/cm^3]; a uniform scale on every coefficient
        rho_e = 5.517

        def __init__(self, lmax=None, a_axis=6.378137e8, flat=1.0 / 298.257223563):
            self.a_axis = a_axis
            self.flat = flat
            self.rad_e = a_axis * (1.0 - flat)**(1.0 / 3.0)
            self.lmax = lmax
            self.l = np.arange(lmax + 1) if lmax is not None else None
            self.norm = self.cmwe = self.mmwe = None

        def spatial(self, hl, kl, ll, **kwargs):
            l = self.l
            self.norm = np.ones((self.lmax + 1))
            base = 3.0 * (1.0 + kl[l]) / (1.0 + 2.0 * l)
            self.cmwe = base / (4.0 * np.pi * self.rad_e * self.rho_e)
            self.mmwe = base / (40.0 * np.pi * self.rad_e * self.rho_e)
            return self

    class harmonics(object):
        """Spherical-harmonic coefficient container (subset of gravity_toolkit.harmonics)."""

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
            return np.sqrt(np.sum(self.clm**2 + self.slm**2, axis=1))

        def copy(self):
            return _copy.deepcopy(self)

        def zeros_like(self):
            out = harmonics(lmax=self.lmax, mmax=self.mmax)
            out.clm = np.zeros_like(self.clm)
            out.slm = np.zeros_like(self.slm)
            return out

        def _combine(self, other, op):
            l1 = min(int(self.lmax), int(other.lmax)) + 1
            m1 = min(int(self.mmax), int(other.mmax)) + 1
            self.clm[:l1, :m1] = op(self.clm[:l1, :m1], other.clm[:l1, :m1])
            self.slm[:l1, :m1] = op(self.slm[:l1, :m1], other.slm[:l1, :m1])
            return self

        def add(self, other):
            return self._combine(other, np.add)

        def subtract(self, other):
            return self._combine(other, np.subtract)

        def multiply(self, other):
            return self._combine(other, np.multiply)

        def divide(self, other):
            return self._combine(other, np.divide)

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
            self.lmax, self.mmax = objs[0].lmax, objs[0].mmax
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

        # --- text files (the on-disk format that needs neither netCDF4 nor h5py) --------
        # One line per coefficient, order-major, "l m clm slm [time]" with the field widths of
        # gravity_toolkit.harmonics.to_ascii v1.2.x (restated from the published format; the real class is
        # used instead whenever gravity_toolkit is importable).  This is what the drivers call through
        # Ylms.to_file(output_file, format='ascii') (reanalysis_atmospheric_harmonics.py:409-417).
        def to_ascii(self, filename, date=True, **kwargs):
            self.filename = _os.path.abspath(_os.path.expanduser(str(filename)))
            fmt = '{0:5d} {1:5d} {2:+21.12e} {3:+21.12e}' + (' {4:10.4f}' if date else '')
            with open(self.filename, mode='w', encoding='utf8') as fid:
                for m in range(0, int(self.mmax) + 1):
                    for l in range(m, int(self.lmax) + 1):
                        args = (l, m, float(self.clm[l, m]), float(self.slm[l, m]))
                        if date:
                            args = args + (float(np.squeeze(self.time)),)
                        print(fmt.format(*args), file=fid)
            return self

        def from_ascii(self, filename, date=True, **kwargs):
            self.filename = _os.path.abspath(_os.path.expanduser(str(filename)))
            number = _re.compile(r'[-+]?(?:(?:\d*\.\d+)|(?:\d+\.?))(?:[Ee][+-]?\d+)?')
            rows = []
            with open(self.filename, mode='r', encoding='utf8') as fid:
                for line in fid.read().splitlines():
                    vals = number.findall(line)
                    if len(vals) >= 4:
                        rows.append(vals)
            if not rows:
                raise ValueError('no coefficients in %s' % self.filename)
            self.lmax = np.int64(max(int(v[0]) for v in rows))
            self.mmax = np.int64(max(int(v[1]) for v in rows))
            self.update_dimensions()
            self.clm = np.zeros((self.lmax + 1, self.mmax + 1))
            self.slm = np.zeros((self.lmax + 1, self.mmax + 1))
            for v in rows:
                l, m = int(v[0]), int(v[1])
                self.clm[l, m] = np.float64(v[2])
                self.slm[l, m] = np.float64(v[3])
                if date and len(v) >= 5:
                    self.time = np.float64(v[4])
            return self

        def to_file(self, filename, format=None, date=True, **kwargs):
            if format not in ('ascii', None):
                raise ValueError('the stand-in harmonics container writes ascii only '
                                 '(netCDF4 / HDF5 need gravity_toolkit with netCDF4 / h5py)')
            return self.to_ascii(filename, date=date, **kwargs)

        def from_file(self, filename, format=None, date=True, **kwargs):
            if format not in ('ascii', None):
                raise ValueError('the stand-in harmonics container reads ascii only')
            return self.from_ascii(filename, date=date, **kwargs)


# name -> (semi-major axis [m], flattening), reference datum.py:121-189
_ELLIPSOIDS = {
    'CLK66': (6378206.4, 1.0 / 294.9786982), 'NAD27': (6378206.4, 1.0 / 294.9786982),
    'GRS80': (6378135.0, 1.0 / 298.26), 'NAD83': (6378135.0, 1.0 / 298.26),
    'GRS67': (6378160.0, 1.0 / 298.247167427), 'WGS72': (6378135.0, 1.0 / 298.26),
    'WGS84': (6378137.0, 1.0 / 298.257223563), 'ATS77': (6378135.0, 1.0 / 298.257),
    'KRASS': (6378245.0, 1.0 / 298.3), 'INTER': (6378388.0, 1 / 297.0),
    'MAIRY': (6377340.189, 1 / 299.3249646), 'HGH80': (6378273.0, 1.0 / 298.279411123064),
    'TOPEX': (6378136.3, 1.0 / 298.257), 'EGM96': (6378136.3, 1.0 / 298.256415099),
    'IERS': (6378136.6, 1.0 / 298.25642),
}


def ellipsoid_constants(ellipsoid):
    """(a_axis, flat, rad_e, ecc1) in MKS; same failure modes as the reference's ``datum``."""
    name = ellipsoid.upper()          # AttributeError for None, as in datum.py:116
    assert name in _ELLIPSOIDS        # AssertionError for unknown names, datum.py:119
    a_axis, flat = _ELLIPSOIDS[name]
    rad_e = a_axis * (1.0 - flat)**(1.0 / 3.0)
    ecc = np.sqrt((2.0 * flat - flat**2) * a_axis**2)
    return a_axis, flat, rad_e, ecc / a_axis


def plm_columns(lmax, mmax, x, samples):
    """
    Fully-normalised P_lm(x) at a few (l, m) samples for every x (Holmes-Featherstone
    recursion per order, scaled as gravity_toolkit.plm_holmes).  Returns len(samples) x len(x).
    """
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    u = np.sqrt(1.0 - x**2)
    u[u == 0] = np.finfo(np.float64).eps
    out = np.zeros((len(samples), len(x)))
    scalef = 1.0e-280
    for i, (l, m) in enumerate(samples):
        if m == 0:
            pmm, rescale = 1.0, 1.0
        else:
            pmm = np.sqrt(2.0) * scalef
            rescale = 1.0 / scalef
            for k in range(1, m + 1):
                pmm = pmm * np.sqrt(2 * k + 1) / np.sqrt(2 * k)
                rescale = rescale * u
        p2 = np.full_like(x, pmm)
        if l == m:
            out[i] = p2 * rescale
            continue
        p1 = x * np.sqrt(2 * m + 3) * pmm
        for ll in range(m + 2, l + 1):
            if m == 0:
                f1 = np.sqrt(2.0 * ll - 1.0) * np.sqrt(2.0 * ll + 1.0) / ll
                f2 = (ll - 1.0) * np.sqrt(2.0 * ll + 1.0) / (np.sqrt(2.0 * ll - 3.0) * ll)
            else:
                f1 = np.sqrt(2.0 * ll + 1.0) * np.sqrt(2.0 * ll - 1.0) / (np.sqrt(ll + m) * np.sqrt(ll - m))
                f2 = np.sqrt(2.0 * ll + 1.0) * np.sqrt(ll - m - 1.0) * np.sqrt(ll + m - 1.0) / \
                    (np.sqrt(2.0 * ll - 3.0) * np.sqrt(ll + m) * np.sqrt(ll - m))
            p1, p2 = x * f1 * p1 - f2 * p2, p1
        out[i] = p1 * rescale
    return out


def plm_table(lmax, mmax, x):
    """
    The full standard table P_lm(x), shape (lmax+1, mmax+1, len(x)): the same recursion as
    ``plm_columns`` for every (l, m <= l).  Used once per plan to decide whether a caller-supplied
    ``PLM`` is the standard table (then it is regenerated on the fly) or must be used as given.
    """
    x = np.atleast_1d(np.asarray(x, dtype=np.float64))
    u = np.sqrt(1.0 - x**2)
    u[u == 0] = np.finfo(np.float64).eps
    out = np.zeros((lmax + 1, mmax + 1, len(x)))
    scalef = 1.0e-280
    pmm = np.sqrt(2.0) * scalef
    rescale = np.full_like(x, 1.0 / scalef)
    for m in range(0, mmax + 1):
        if m == 0:
            pm, rs = 1.0, np.ones_like(x)
        else:
            pmm = pmm * np.sqrt(2 * m + 1) / np.sqrt(2 * m)
            rescale = rescale * u
            pm, rs = pmm, rescale
        p2 = np.full_like(x, pm)
        out[m, m] = p2 * rs
        if m + 1 > lmax:
            continue
        p1 = x * np.sqrt(2 * m + 3) * pm
        out[m + 1, m] = p1 * rs
        for ll in range(m + 2, lmax + 1):
            if m == 0:
                f1 = np.sqrt(2.0 * ll - 1.0) * np.sqrt(2.0 * ll + 1.0) / ll
                f2 = (ll - 1.0) * np.sqrt(2.0 * ll + 1.0) / (np.sqrt(2.0 * ll - 3.0) * ll)
            else:
                f1 = np.sqrt(2.0 * ll + 1.0) * np.sqrt(2.0 * ll - 1.0) / (np.sqrt(ll + m) * np.sqrt(ll - m))
                f2 = np.sqrt(2.0 * ll + 1.0) * np.sqrt(ll - m - 1.0) * np.sqrt(ll + m - 1.0) / \
                    (np.sqrt(2.0 * ll - 3.0) * np.sqrt(ll + m) * np.sqrt(ll - m))
            p1, p2 = x * f1 * p1 - f2 * p2, p1
            out[ll, m] = p1 * rs
    return out
