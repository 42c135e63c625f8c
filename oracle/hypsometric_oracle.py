"""
NumPy restatement of the upstream producer of (geopotential height, pressure difference) cubes
(TEST INFRASTRUCTURE -- see ``oracle/__init__.py``): the hypsometric integration over model levels
of reference ``model_harmonics/reanalysis/reanalysis_geopotential_heights.py:349-383``.

The reference code is inline in a netCDF driver, so it cannot be called; ``reference_loop`` below
executes those very source lines, read from ``/root/reference`` at run time, in a namespace of
synthetic arrays.  ``tests/test_oracle_hypsometric.py`` checks the restatement against it bit for bit
and against ``tests/golden/hypsometric.npz`` (which travels to the GPU box).

Mixed precision of the reference (NumPy 2.3.3 pinned in its pixi.lock; same rules as 2.3.5 here),
reproduced here and in the kernel: temperature and humidity arrive as float32 but are wrapped in
masked arrays (:346-347), and masked-array arithmetic promotes the Python scalars strongly, so the
virtual temperature ``(1 + 0.609133 q) T`` and ``R_dry * Tv`` are float64 operations on the widened
values; the level pressures ``A + B*SP`` and the logarithm are float64; the running geopotential is a
float32 array updated in place (``float32(float64(gh) + term)``), and both outputs are stored as
float32.
"""
import os
import textwrap

import numpy as np

from oracle.ref_loader import REFERENCE_ROOT

_REF_FILE = os.path.join(REFERENCE_ROOT, 'model_harmonics', 'reanalysis', 'reanalysis_geopotential_heights.py')
_REF_LINES = (349, 383)       # 1-based, inclusive


def geopotential_levels(T, Q, SP, geopotential, A, B, AI, BI, GRAVITY=1.0, R_dry=287.06, Pmin=0.1):
    """
    T, Q: (nlevels, nlat, nlon), bottom level first; SP, geopotential: (nlat, nlon); A, B: half-level
    coefficients; AI, BI: interface coefficients.  Returns (Z, DIFF) float32 (nlevels, nlat, nlon):
    geopotential at the levels [m^2/s^2] and pressure difference across them [Pa].
    Follows reanalysis_geopotential_heights.py:349-383.
    """
    nlevels, nlat, nlon = np.shape(T)
    Z = np.zeros((nlevels, nlat, nlon), dtype='f')
    DIFF = np.zeros((nlevels, nlat, nlon), dtype='f')
    gh = np.empty((nlat, nlon), dtype=np.float32)
    gh[:, :] = geopotential * GRAVITY                                    # :349-351
    for k in range(nlevels):                                             # :355
        # :360 -- float64: the reference's operands are masked arrays (see the module docstring)
        Tv = (1.0 + 0.609133 * Q[k, :, :].astype(np.float64)) * T[k, :, :].astype(np.float64)
        Pnum = A[k] + B[k] * SP                                          # :362
        if (k + 1) == len(A):                                            # :363-368
            Pdom = Pmin
        else:
            Pdom = np.maximum(A[k + 1] + B[k + 1] * SP, Pmin)
        gh[:, :] += R_dry * Tv * np.log(Pnum / Pdom)                     # :371
        Z[k, :, :] = gh                                                  # :373
        Plower = AI[k] + BI[k] * SP                                      # :375
        if (k + 1) == len(AI):                                           # :377-382
            Pupper = Pmin
        else:
            Pupper = np.maximum(AI[k + 1] + BI[k + 1] * SP, Pmin)
        DIFF[k, :, :] = Pupper - Plower                                  # :383
    return Z, DIFF


def reference_available():
    return os.path.isfile(_REF_FILE)


def reference_loop(T, Q, SP, geopotential, A, B, AI, BI, GRAVITY=1.0, R_dry=287.06, Pmin=0.1):
    """Runs lines 349-383 of the reference file, unmodified, on the given arrays."""
    src = open(_REF_FILE).read().splitlines()[_REF_LINES[0] - 1:_REF_LINES[1]]
    code = textwrap.dedent('\n'.join(src))
    nlevels, nlat, nlon = np.shape(T)
    ns = dict(np=np, t_time=np.ma.masked_equal(T, -9999.0), q_time=np.ma.masked_equal(Q, -9999.0), nlat=nlat,
              nlon=nlon, nlevels=nlevels, geopotential=geopotential, GRAVITY=GRAVITY, pressure=SP[np.newaxis], t=0,
              A=A, B=B, AI=AI, BI=BI, Pmin=Pmin, R_dry=R_dry, ZNAME='z', DIFFNAME='dp',
              dinput={'z': np.ma.zeros((1, nlevels, nlat, nlon), dtype='f'),
                      'dp': np.ma.zeros((1, nlevels, nlat, nlon), dtype='f')})
    exec(compile(code, _REF_FILE, 'exec'), ns)
    return np.ma.getdata(ns['dinput']['z'][0]), np.ma.getdata(ns['dinput']['dp'][0])


def synthetic_model_levels(nlev, nlat, nlon, seed=1234, interfaces_extra=True):
    """Seeded synthetic hybrid-level inputs (shared generator in model_harmonics_b200.testing)."""
    from model_harmonics_b200.testing import model_level_inputs
    return model_level_inputs(nlev, nlat, nlon, seed=seed, interfaces_extra=interfaces_extra)
