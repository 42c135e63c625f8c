"""
Upstream producer of ``gen_atmosphere_stokes``' inputs: geopotential at hybrid
model levels and the pressure differences across them, from temperature, specific humidity and
surface pressure -- the arithmetic of reference
``model_harmonics/reanalysis/reanalysis_geopotential_heights.py:349-383`` (there inline in a netCDF
driver; here a function), computed by ``csrc/hypsometric.cu``.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .plan import require_device

__all__ = ['geopotential_levels']


def _f32_dev(a, device, shape=None):
    if isinstance(a, torch.Tensor):
        t = a.to('cuda:%d' % device, non_blocking=True)
    else:
        a = np.ma.getdata(a) if isinstance(a, np.ma.MaskedArray) else np.asarray(a)
        if a.dtype != np.float32:
            raise TypeError('temperature, humidity, surface pressure and geopotential must be float32 '
                            '(their on-disk type); got %s' % a.dtype)
        t = torch.from_numpy(np.ascontiguousarray(a)).to('cuda:%d' % device, non_blocking=True)
    if t.dtype != torch.float32:
        raise TypeError('expected float32 operands')
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError('operand shape %s does not match %s' % (tuple(t.shape), tuple(shape)))
    return t.contiguous()


def geopotential_levels(T, Q, SP, geopotential, A, B, AI, BI, GRAVITY=1.0, R_dry=287.06, Pmin=0.1,
                        divide_by=1.0, as_numpy=True):
    """
    T, Q: float32 (nlevels, nlat, nlon), bottom level first; SP: float32 (nlat, nlon) surface
    pressure; ``geopotential``: float32 (nlat, nlon) surface geopotential (multiplied by ``GRAVITY``
    as the reference does); A, B / AI, BI: float64 hybrid coefficients at half levels / interfaces.
    Returns float32 ``(Z / divide_by, DIFF)`` of shape (nlevels, nlat, nlon); with
    ``as_numpy=False`` they stay on the device and can be passed straight to
    ``gen_atmosphere_stokes`` / ``atmosphere_stokes_batch`` (``divide_by=9.80665`` gives the
    geopotential height the harmonics driver forms at ``reanalysis_atmospheric_harmonics.py:357``).
    """
    device = require_device()
    nlev, nlat, nlon = tuple(T.shape)
    t = _f32_dev(T, device)
    q = _f32_dev(Q, device, (nlev, nlat, nlon))
    sp = _f32_dev(SP, device, (nlat, nlon))
    if isinstance(geopotential, torch.Tensor) and geopotential.is_cuda:
        # float32 array times a Python scalar stays float32 in NumPy; same single rounding here
        gh0_t = (geopotential.to(torch.float32) * float(GRAVITY)).contiguous() if GRAVITY != 1.0 \
            else geopotential.to(torch.float32).contiguous()
    else:
        gh0 = np.empty((nlat, nlon), dtype=np.float32)
        g = geopotential.cpu().numpy() if isinstance(geopotential, torch.Tensor) else np.asarray(geopotential)
        gh0[:, :] = g * GRAVITY                                          # :349-351
        gh0_t = torch.from_numpy(gh0).to('cuda:%d' % device)
    if tuple(gh0_t.shape) != (nlat, nlon):
        raise ValueError('geopotential must be (nlat, nlon)')
    A, B, AI, BI = [np.ascontiguousarray(np.ma.getdata(v), dtype=np.float64) for v in (A, B, AI, BI)]
    if len(A) != len(B) or len(AI) != len(BI):
        raise ValueError('hybrid coefficient vectors must pair up')
    Z = torch.empty((nlev, nlat, nlon), dtype=torch.float32, device=t.device)
    D = torch.empty_like(Z)
    rc = _lib.lib().mhb200_geopotential_levels_f32(
        device, ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(q.data_ptr()), ctypes.c_void_p(sp.data_ptr()),
        ctypes.c_void_p(gh0_t.data_ptr()), _lib.as_double_p(A), _lib.as_double_p(B), len(A), _lib.as_double_p(AI),
        _lib.as_double_p(BI), len(AI), int(nlev), int(nlat * nlon), float(R_dry), float(Pmin), float(divide_by),
        ctypes.c_void_p(Z.data_ptr()), ctypes.c_void_p(D.data_ptr()),
        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _lib.check(rc, 'mhb200_geopotential_levels_f32')
    return (Z.cpu().numpy(), D.cpu().numpy()) if as_numpy else (Z, D)
