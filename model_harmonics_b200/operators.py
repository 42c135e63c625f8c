"""
Drop-in operators: same names, positional order, keyword names, defaults and return type as
the reference's ``gen_pressure_stokes`` / ``gen_atmosphere_stokes`` / ``gen_point_pressure``
(``model_harmonics/__init__.py:19-21``), computed by the sm_100a kernels in ``csrc/``.

The host code here only does the O(nlat)+O(nlon)+O(L) preparation, expression for expression
as the reference does it, and hands device pointers to the C ABI (``include/mhb200.h``).
Array operands may be NumPy arrays (copied to the device) or torch CUDA tensors (zero copy).
"""
import ctypes
import threading

import numpy as np
import torch

from . import _lib
from . import _compat
from .plan import get_plan, plm_table_for, require_device

__all__ = ['gen_pressure_stokes', 'gen_atmosphere_stokes', 'gen_point_pressure', 'gen_stokes',
           'pressure_stokes_batch', 'atmosphere_stokes_batch', 'point_pressure_batch', 'stokes_batch',
           'gen_point_pressure_tiles']


def _stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _to_device(a, device, dtype=torch.float64, keep_f32=False):
    """NumPy / scalar / torch operand -> device tensor (no copy for CUDA tensors)."""
    if isinstance(a, torch.Tensor):
        t = a if a.is_cuda else a.to('cuda:%d' % device, non_blocking=True)
    else:
        if isinstance(a, np.ma.MaskedArray):
            a = np.ma.getdata(a)
        arr = np.asarray(a)
        if arr.dtype == np.float32 and keep_f32:
            pass
        elif arr.dtype != np.float64:
            arr = arr.astype(np.float64)
        # torch.from_numpy rejects negative strides (lat-flipped views such as P[::-1, :], which the
        # reference accepts); positive-stride views of 2-D+ operands stay views (addressed with strides)
        if (not arr.flags.c_contiguous and arr.ndim < 2) or any(st < 0 for st in arr.strides):
            arr = np.ascontiguousarray(arr)
        t = torch.from_numpy(np.atleast_1d(arr)).to('cuda:%d' % device, non_blocking=True)
    if keep_f32 and t.dtype == torch.float32:
        return t
    if t.dtype != dtype:
        t = t.to(dtype)
    return t


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _make_harmonics(clm, slm, lmax, mmax):
    Ylms = _compat.harmonics(lmax=lmax, mmax=mmax)
    Ylms.clm = clm
    Ylms.slm = slm
    return Ylms


def _weights(phi, th, WEIGHT):
    # reference gen_pressure_stokes.py:163-172
    nlat = len(th)
    int_fact = np.zeros((nlat))
    if WEIGHT is not None:
        int_fact[:] = np.broadcast_to(np.atleast_1d(WEIGHT), nlat)
    else:
        dphi = np.abs(phi[1] - phi[0])
        dth = np.abs(th[1] - th[0])
        int_fact[:] = np.sin(th) * dphi * dth
    return int_fact


def _grid_operand(a, device, nlat, nlon, lat_first, name):
    """Device tensor and (field, lat, lon) element strides for P-like operands."""
    t = _to_device(a, device)
    if t.numel() == 1:
        return t.reshape(1), (0, 0, 0)
    if t.dim() == 1:
        if t.shape[0] != nlat:
            raise ValueError('%s: 1-D operand must have one value per latitude' % name)
        return t, (0, t.stride(0), 0)
    if t.dim() != 2:
        raise ValueError('%s: expected a scalar, 1-D or 2-D operand' % name)
    want = (nlat, nlon) if lat_first else (nlon, nlat)
    if tuple(t.shape) != want:
        raise ValueError('%s: shape %s does not match the grid %s' % (name, tuple(t.shape), want))
    s0, s1 = t.stride()
    return t, ((0, s0, s1) if lat_first else (0, s1, s0))


# -----------------------------------------------------------------------------------------
# gridded surface pressure
# -----------------------------------------------------------------------------------------
def _pressure_core(plan, P_t, Ps, G_t, Gs, R_t, Rs, nbatch, rad_e, dfactor, plm_tab):
    L = _lib.lib()
    dev = 'cuda:%d' % plan.device
    clm = torch.empty((nbatch, plan.lmax + 1, plan.mmax + 1), dtype=torch.float64, device=dev)
    slm = torch.empty_like(clm)
    ws = plan.workspace(nbatch)
    dfactor = np.ascontiguousarray(dfactor, dtype=np.float64)
    with plan.lock:
        rc = L.mhb200_pressure_stokes_f64(plan.handle, _ptr(P_t), _lib.int64x(*Ps), _ptr(G_t), _lib.int64x(*Gs),
                                          _ptr(R_t), _lib.int64x(*Rs), int(nbatch), float(rad_e),
                                          _lib.as_double_p(dfactor), _ptr(plm_tab), _ptr(clm), _ptr(slm),
                                          _ptr(ws), ws.numel(), _stream_ptr())
    _lib.check(rc, 'mhb200_pressure_stokes_f64')
    return clm, slm


def gen_pressure_stokes(P, G, R, lon, lat, LMAX=60, MMAX=None, WEIGHT=None, PLM=None, LOVE=None):
    r"""
    Converts a surface pressure field to fully-normalised Stokes coefficients
    (reference ``model_harmonics/gen_pressure_stokes.py:81-209``; same arguments).

    Returns a ``harmonics`` object with ``clm``/``slm`` of shape ``(LMAX+1, MMAX+1)``.
    """
    device = require_device()
    LMAX = np.int64(LMAX)
    MMAX = np.copy(LMAX) if not MMAX else MMAX                       # :133-135
    phi = np.radians(np.squeeze(lon))                                # :138-141
    th = np.radians(90.0 - np.squeeze(lat))
    phi = np.where(phi < 0, phi + 2.0 * np.pi, phi)
    nlat, nlon = len(th), len(phi)
    lat_first = (np.shape(P)[0] == nlat)                             # :146-151
    factors = _compat.units(lmax=LMAX).spatial(*LOVE)                # :155-160
    rad_e = factors.rad_e / 100.0
    dfactor = factors.mmwe
    int_fact = _weights(phi, th, WEIGHT)
    plan = get_plan(phi, np.cos(th), int_fact, int(LMAX), int(MMAX), device)
    plm_tab = plm_table_for(plan, PLM, int(LMAX), int(MMAX))
    P_t, Ps = _grid_operand(P, device, nlat, nlon, lat_first, 'P')
    if P_t.dim() != 2:
        raise ValueError('P must be a 2-D grid')
    G_t, Gs = _grid_operand(G, device, nlat, nlon, lat_first, 'G')
    R_t, Rs = _grid_operand(R, device, nlat, nlon, lat_first, 'R')
    clm, slm = _pressure_core(plan, P_t, Ps, G_t, Gs, R_t, Rs, 1, rad_e, dfactor, plm_tab)
    return _make_harmonics(clm[0].cpu().numpy(), slm[0].cpu().numpy(), LMAX, MMAX)


def pressure_stokes_batch(P, G, R, lon, lat, LMAX=60, MMAX=None, WEIGHT=None, LOVE=None, as_numpy=True):
    """
    Batched form for the driver pattern "static G, R, varying P" (reference
    ``reanalysis_pressure_harmonics.py:472-503``): ``P`` is (T, nlat, nlon); G and R are grids
    shared by all fields (or scalars).  Returns (clm, slm) of shape (T, LMAX+1, MMAX+1).
    """
    device = require_device()
    LMAX = np.int64(LMAX)
    MMAX = np.copy(LMAX) if not MMAX else MMAX
    phi = np.radians(np.squeeze(lon))
    th = np.radians(90.0 - np.squeeze(lat))
    phi = np.where(phi < 0, phi + 2.0 * np.pi, phi)
    nlat, nlon = len(th), len(phi)
    factors = _compat.units(lmax=LMAX).spatial(*LOVE)
    plan = get_plan(phi, np.cos(th), _weights(phi, th, WEIGHT), int(LMAX), int(MMAX), device)
    P_t = _to_device(P, device)
    if P_t.dim() != 3 or tuple(P_t.shape[1:]) != (nlat, nlon):
        raise ValueError('P must be (T, nlat, nlon)')
    G_t, Gs = _grid_operand(G, device, nlat, nlon, True, 'G')
    R_t, Rs = _grid_operand(R, device, nlat, nlon, True, 'R')
    clm, slm = _pressure_core(plan, P_t, P_t.stride(), G_t, Gs, R_t, Rs, P_t.shape[0],
                              factors.rad_e / 100.0, factors.mmwe, None)
    return (clm.cpu().numpy(), slm.cpu().numpy()) if as_numpy else (clm, slm)


# -----------------------------------------------------------------------------------------
# radius-free sibling: gravity_toolkit.gen_stokes
# -----------------------------------------------------------------------------------------
def _stokes_prepare(lon, lat, LMAX, MMAX, UNITS, LOVE, WEIGHT, device):
    LMAX = np.int64(LMAX)
    MMAX = np.copy(LMAX) if (MMAX is None) else MMAX
    lon, lat = np.squeeze(lon), np.squeeze(lat)
    phi = lon * np.pi / 180.0
    th = (90.0 - lat) * np.pi / 180.0
    factors = _compat.units(lmax=LMAX).spatial(*LOVE)
    int_fact = _weights(phi, th, WEIGHT)
    if isinstance(UNITS, (list, tuple, np.ndarray)):
        dfactor = np.array(UNITS, dtype=np.float64)              # custom degree-dependent factors
    elif UNITS == 1:
        dfactor = factors.cmwe                                   # cm w.e. (g/cm^2)
    elif UNITS == 2:
        dfactor = factors.cmwe                                   # Gt: cell masses, no area weighting
        int_fact = np.full(len(th), 1.0e15 / (factors.rad_e**2))
    elif UNITS == 3:
        dfactor = factors.mmwe                                   # kg/m^2 (mm w.e.)
    else:
        raise ValueError('Unknown units %r' % (UNITS,))
    plan = get_plan(phi, np.cos(th), int_fact, int(LMAX), int(MMAX), device)
    return LMAX, MMAX, plan, np.asarray(dfactor, dtype=np.float64), len(th), len(phi)


def gen_stokes(data, lon, lat, LMIN=0, LMAX=60, MMAX=None, UNITS=1, PLM=None, LOVE=None, WEIGHT=None):
    """
    Spherical harmonics of a gridded surface-mass field with no radial factor: the sibling operator the
    2-D reanalysis driver and the SMB/TWS drivers call (``gravtk.gen_stokes``, reference call sites
    ``reanalysis/reanalysis_monthly_harmonics.py:434``, ``TWS/gldas_monthly_harmonics.py:373``,
    ``SMB/era5_smb_harmonics.py:248``).  It is ``gen_pressure_stokes`` with ``(R/rad_e)^(l+2)`` and the
    gravity division removed, so it runs on the same kernels with R = rad_e = G = 1.  ``data`` may be
    (lat, lon) or (lon, lat); ``UNITS``: 1 cm w.e., 2 Gt, 3 kg/m^2, or an array of degree factors;
    ``WEIGHT`` (latitude integration weights) as the reanalysis driver passes it.
    The third-party function is restated from its published algorithm (not vendored under /root/reference).
    """
    device = require_device()
    LMAX, MMAX, plan, dfactor, nlat, nlon = _stokes_prepare(lon, lat, LMAX, MMAX, UNITS, LOVE, WEIGHT, device)
    plm_tab = plm_table_for(plan, PLM, int(LMAX), int(MMAX))
    lat_first = (np.shape(data)[0] == nlat)
    P_t, Ps = _grid_operand(data, device, nlat, nlon, lat_first, 'data')
    if P_t.dim() != 2:
        raise ValueError('data must be a 2-D grid')
    one = torch.ones(1, dtype=torch.float64, device='cuda:%d' % device)
    clm, slm = _pressure_core(plan, P_t, Ps, one, (0, 0, 0), one, (0, 0, 0), 1, 1.0, dfactor, plm_tab)
    clm, slm = clm[0].cpu().numpy(), slm[0].cpu().numpy()
    if LMIN > 0:
        clm[:int(LMIN), :] = 0.0
        slm[:int(LMIN), :] = 0.0
    return _make_harmonics(clm, slm, LMAX, MMAX)


def stokes_batch(data, lon, lat, LMAX=60, MMAX=None, UNITS=1, LOVE=None, WEIGHT=None, as_numpy=True):
    """Batched ``gen_stokes``: ``data`` is (T, nlat, nlon); returns (clm, slm) of shape (T, LMAX+1, MMAX+1)."""
    device = require_device()
    LMAX, MMAX, plan, dfactor, nlat, nlon = _stokes_prepare(lon, lat, LMAX, MMAX, UNITS, LOVE, WEIGHT, device)
    P_t = _to_device(data, device)
    if P_t.dim() != 3 or tuple(P_t.shape[1:]) != (nlat, nlon):
        raise ValueError('data must be (T, nlat, nlon)')
    one = torch.ones(1, dtype=torch.float64, device='cuda:%d' % device)
    clm, slm = _pressure_core(plan, P_t, P_t.stride(), one, (0, 0, 0), one, (0, 0, 0), P_t.shape[0], 1.0, dfactor,
                              None)
    return (clm.cpu().numpy(), slm.cpu().numpy()) if as_numpy else (clm, slm)


# -----------------------------------------------------------------------------------------
# 3-D atmosphere
# -----------------------------------------------------------------------------------------
def _atmosphere_tables(plan, th, phi, ellipsoid):
    key = ('atm', ellipsoid.upper())
    with plan.lock:
        if key in plan.tables:
            return plan.tables[key]
    a_axis, flat, rad_e, ecc1 = _compat.ellipsoid_constants(ellipsoid)
    # colatitude-only factors, formed as the reference forms them on the meshgrid
    # (gen_atmosphere_stokes.py:171,211-215,222-224,232-238)
    N = a_axis / np.sqrt(1.0 - ecc1**2.0 * np.cos(th)**2.0)
    ge = 9.780356
    ring = np.zeros((9, len(th)))
    ring[0] = (1.0 - 0.002644 * np.cos(2.0 * th))
    ring[1] = (1.0 - 0.0089 * np.cos(2.0 * th)) / 6.245e6
    ring[2] = N
    ring[3] = N * (1.0 - ecc1**2.0)
    ring[4] = np.sin(th)
    ring[5] = np.cos(th)
    ring[6] = ge * (1.0 + 5.2885e-3 * np.cos(th)**2 - 5.9e-6 * np.cos(2.0 * th)**2)
    ring[7] = 2.0 * (1.006803 - 0.06706 * np.cos(th)**2)
    lon_tab = np.stack([np.cos(phi), np.sin(phi)])
    dev = 'cuda:%d' % plan.device
    out = (torch.from_numpy(ring).to(dev), torch.from_numpy(np.ascontiguousarray(lon_tab)).to(dev))
    with plan.lock:
        plan.tables[key] = out
    return out


def _atmosphere_core(plan, GPH_t, dP_t, batch_stride, field_scale, geoid_t, geoid_strides, ring_t, lon_t,
                     nlev, nbatch, method, rad_e, dfactor, plm_tab):
    L = _lib.lib()
    dev = 'cuda:%d' % plan.device
    clm = torch.empty((nbatch, plan.lmax + 1, plan.mmax + 1), dtype=torch.float64, device=dev)
    slm = torch.empty_like(clm)
    ws = plan.workspace(nbatch, nlev)
    dfactor = np.ascontiguousarray(dfactor, dtype=np.float64)
    dtype = _lib.F32 if GPH_t.dtype == torch.float32 else _lib.F64
    fs = None
    if field_scale is not None:
        fs = np.ascontiguousarray(field_scale, dtype=np.float64).reshape(-1)
        if fs.size != 2 * nbatch:
            raise ValueError('field_scale must hold 2 values per field')
    with plan.lock:
        rc = L.mhb200_atmosphere_stokes(plan.handle, dtype, _ptr(GPH_t), _ptr(dP_t), int(batch_stride),
                                        _lib.as_double_p(fs) if fs is not None else None,
                                        _ptr(geoid_t), _lib.int64x(*geoid_strides), _ptr(ring_t), _ptr(lon_t),
                                        int(nlev), int(nbatch), int(method), float(rad_e),
                                        _lib.as_double_p(dfactor), _ptr(plm_tab), _ptr(clm), _ptr(slm),
                                        _ptr(ws), ws.numel(), _stream_ptr())
    _lib.check(rc, 'mhb200_atmosphere_stokes')
    return clm, slm


def _method_code(METHOD):
    if METHOD == 'BC05':
        return _lib.METHOD_BC05
    if METHOD == 'SW02':
        return _lib.METHOD_SW02
    # the reference silently returns zeros for an unknown METHOD (no else branch at
    # gen_atmosphere_stokes.py:254-265); raising is this package's one deliberate deviation
    raise ValueError("METHOD must be 'BC05' or 'SW02'")


def _atmosphere_prepare(GPH_shape, lon, lat, LMAX, MMAX, ELLIPSOID, GEOID, WEIGHT, PLM, LOVE, device,
                        host_geoid_ok=False):
    LMAX = np.int64(LMAX)
    MMAX = np.copy(LMAX) if not MMAX else MMAX                       # :147-149
    nlat, nlon = GPH_shape[-2], GPH_shape[-1]
    phi = np.radians(np.squeeze(lon))                                # :154-155 (no 2*pi wrap here)
    th = np.radians(90.0 - np.squeeze(lat))
    if len(th) != nlat or len(phi) != nlon:
        raise ValueError('GPH must be (level, lat, lon)')
    a_axis, flat, rad_e, ecc1 = _compat.ellipsoid_constants(ELLIPSOID)   # :160-168
    dfactor = _compat.units(lmax=LMAX, a_axis=100.0 * a_axis, flat=flat).spatial(*LOVE).mmwe   # :175-176
    int_fact = _weights(phi, th, WEIGHT)
    plan = get_plan(phi, np.cos(th), int_fact, int(LMAX), int(MMAX), device)
    ring_t, lon_t = _atmosphere_tables(plan, th, phi, ELLIPSOID)
    plm_tab = plm_table_for(plan, PLM, int(LMAX), int(MMAX))
    if GEOID is None:
        raise TypeError('GEOID is required (geoid height on the (lat, lon) grid, or a scalar)')
    if host_geoid_ok and not isinstance(GEOID, torch.Tensor) and np.shape(GEOID) == (nlat, nlon):
        # host-cube pipeline: the library stages the host geoid itself, behind the first band's cubes
        gh = np.ma.getdata(GEOID) if isinstance(GEOID, np.ma.MaskedArray) else np.asarray(GEOID)
        gh = np.ascontiguousarray(gh, dtype=np.float64)
        return LMAX, MMAX, plan, ring_t, lon_t, plm_tab, gh, None, rad_e, dfactor
    geoid_t = _to_device(GEOID, device)
    if geoid_t.numel() == 1:
        geoid_t, gstr = geoid_t.reshape(1), (0, 0)
    elif tuple(geoid_t.shape) == (nlat, nlon):
        gstr = tuple(geoid_t.stride())
    else:
        raise ValueError('GEOID must be a scalar or a (lat, lon) grid')
    return LMAX, MMAX, plan, ring_t, lon_t, plm_tab, geoid_t, gstr, rad_e, dfactor


def _cube_pair(GPH, pressure, device):
    f32 = all((isinstance(a, torch.Tensor) and a.dtype == torch.float32) or
              (not isinstance(a, torch.Tensor) and np.asarray(a).dtype == np.float32) for a in (GPH, pressure))
    g = _to_device(GPH, device, keep_f32=f32)
    d = _to_device(pressure, device, keep_f32=f32)
    if g.shape != d.shape:
        raise ValueError('GPH and pressure must have the same shape')
    return g.contiguous(), d.contiguous()


import os as _os
# latitude bands of the host-cube pipeline (<= 16): 2/4/8/16 bands measured 83.1/83.0/80.9/80.6 fields/s at C3
_HOST_BANDS = int(_os.environ.get('MHB200_HOST_BANDS', '4'))


def _host_cube_pair(GPH, pressure):
    """Both cubes as C-contiguous host arrays of one storage type (float64, or float32 as stored), else None."""
    if isinstance(GPH, torch.Tensor) or isinstance(pressure, torch.Tensor):
        return None
    g = np.ma.getdata(GPH) if isinstance(GPH, np.ma.MaskedArray) else np.asarray(GPH)
    d = np.ma.getdata(pressure) if isinstance(pressure, np.ma.MaskedArray) else np.asarray(pressure)
    if g.shape != d.shape:
        raise ValueError('GPH and pressure must have the same shape')
    if not (g.dtype == np.float32 and d.dtype == np.float32):
        g = g.astype(np.float64, copy=False)
        d = d.astype(np.float64, copy=False)
    return np.ascontiguousarray(g), np.ascontiguousarray(d)


def _atmosphere_host(plan, g, d, geoid_t, geoid_strides, ring_t, lon_t, nlev, method, rad_e, dfactor, plm_tab):
    L = _lib.lib()
    clm = np.empty((plan.lmax + 1, plan.mmax + 1))
    slm = np.empty((plan.lmax + 1, plan.mmax + 1))
    ws = plan.host_workspace(_HOST_BANDS)
    dfactor = np.ascontiguousarray(dfactor, dtype=np.float64)
    dtype = _lib.F32 if g.dtype == np.float32 else _lib.F64
    if geoid_strides is None:
        # geoid_t is a contiguous host grid (see _atmosphere_prepare)
        with plan.lock:
            rc = L.mhb200_atmosphere_stokes_host_geoid(plan.handle, dtype, ctypes.c_void_p(g.ctypes.data),
                                                       ctypes.c_void_p(d.ctypes.data), _lib.as_double_p(geoid_t),
                                                       _ptr(ring_t), _ptr(lon_t), int(nlev), _HOST_BANDS, int(method),
                                                       float(rad_e), _lib.as_double_p(dfactor), _ptr(plm_tab),
                                                       _lib.as_double_p(clm), _lib.as_double_p(slm), _ptr(ws),
                                                       ws.numel(), _stream_ptr())
        _lib.check(rc, 'mhb200_atmosphere_stokes_host_geoid')
        return clm, slm
    with plan.lock:
        rc = L.mhb200_atmosphere_stokes_host(plan.handle, dtype, ctypes.c_void_p(g.ctypes.data),
                                             ctypes.c_void_p(d.ctypes.data), _ptr(geoid_t),
                                             _lib.int64x(*geoid_strides), _ptr(ring_t), _ptr(lon_t), int(nlev),
                                             _HOST_BANDS, int(method), float(rad_e), _lib.as_double_p(dfactor),
                                             _ptr(plm_tab), _lib.as_double_p(clm), _lib.as_double_p(slm),
                                             _ptr(ws), ws.numel(), _stream_ptr())
    _lib.check(rc, 'mhb200_atmosphere_stokes_host')
    return clm, slm


def gen_atmosphere_stokes(GPH, pressure, lon, lat, LMAX=60, MMAX=None, ELLIPSOID=None, GEOID=None,
                          WEIGHT=None, PLM=None, LOVE=None, METHOD='BC05'):
    """
    Converts 3-D geopotential-height and pressure-difference cubes (level, lat, lon) to Stokes
    coefficients (reference ``model_harmonics/gen_atmosphere_stokes.py:86-278``; same arguments).
    float32 cubes are accepted as stored (the on-disk type) and widened exactly on the device.
    """
    device = require_device()
    method = _method_code(METHOD)
    shape = tuple(GPH.shape)
    if len(shape) != 3:
        raise ValueError('GPH must be (level, lat, lon)')
    host = _host_cube_pair(GPH, pressure)
    use_host = host is not None and shape[1] >= 16
    LMAX, MMAX, plan, ring_t, lon_t, plm_tab, geoid_t, gstr, rad_e, dfactor = _atmosphere_prepare(
        shape, lon, lat, LMAX, MMAX, ELLIPSOID, GEOID, WEIGHT, PLM, LOVE, device, host_geoid_ok=use_host)
    if use_host:
        # host cubes: latitude-band pipeline (copy of band b+1 overlaps the kernel of band b)
        clm, slm = _atmosphere_host(plan, host[0], host[1], geoid_t, gstr, ring_t, lon_t, shape[0], method,
                                    rad_e, dfactor, plm_tab)
        return _make_harmonics(clm, slm, LMAX, MMAX)
    g, d = _cube_pair(GPH, pressure, device)
    clm, slm = _atmosphere_core(plan, g, d, 0, None, geoid_t, gstr, ring_t, lon_t, shape[0], 1, method,
                                rad_e, dfactor, plm_tab)
    return _make_harmonics(clm[0].cpu().numpy(), slm[0].cpu().numpy(), LMAX, MMAX)


def atmosphere_stokes_batch(GPH, pressure, lon, lat, LMAX=60, MMAX=None, ELLIPSOID=None, GEOID=None,
                            WEIGHT=None, LOVE=None, METHOD='BC05', field_scale=None, as_numpy=True):
    """
    Batched form: cubes are (T, level, lat, lon), or (level, lat, lon) together with
    ``field_scale`` of shape (T, 2) -- field t is then (s[t,0]*GPH, s[t,1]*pressure), formed on
    load (how the multi-decade benchmark derives 480 fields from one base cube).
    Returns (clm, slm) of shape (T, LMAX+1, MMAX+1).
    """
    device = require_device()
    method = _method_code(METHOD)
    shape = tuple(GPH.shape)
    g, d = _cube_pair(GPH, pressure, device)
    if len(shape) == 4:
        nbatch, nlev = shape[0], shape[1]
        bstride = g.stride(0)
    elif len(shape) == 3 and field_scale is not None:
        nbatch, nlev, bstride = int(np.shape(field_scale)[0]), shape[0], 0
    else:
        raise ValueError('GPH must be (T, level, lat, lon), or (level, lat, lon) with field_scale')
    LMAX, MMAX, plan, ring_t, lon_t, plm_tab, geoid_t, gstr, rad_e, dfactor = _atmosphere_prepare(
        shape, lon, lat, LMAX, MMAX, ELLIPSOID, GEOID, WEIGHT, None, LOVE, device)
    clm, slm = _atmosphere_core(plan, g, d, bstride, field_scale, geoid_t, gstr, ring_t, lon_t, nlev, nbatch,
                                method, rad_e, dfactor, plm_tab)
    return (clm.cpu().numpy(), slm.cpu().numpy()) if as_numpy else (clm, slm)


# -----------------------------------------------------------------------------------------
# scattered points
# -----------------------------------------------------------------------------------------
_point_ws = {}           # (device, CUDA stream handle) -> workspace tensor
_point_ws_lock = threading.Lock()


def _points_core(device, P_t, G_t, Gs, R_t, Rs, phi_t, theta_t, area_t, npts, nbatch, lmax, mmax, rad_e, dfactor):
    L = _lib.lib()
    dev = 'cuda:%d' % device
    need = L.mhb200_points_workspace_bytes(int(npts), int(nbatch), int(lmax), int(mmax))
    if need == 0:
        raise ValueError('bad sizes for the point operator')
    key = (device, int(torch.cuda.current_stream().cuda_stream))
    with _point_ws_lock:
        ws = _point_ws.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
            _point_ws[key] = ws
    clm = torch.empty((nbatch, lmax + 1, mmax + 1), dtype=torch.float64, device=dev)
    slm = torch.empty_like(clm)
    dfactor = np.ascontiguousarray(dfactor, dtype=np.float64)
    rc = L.mhb200_point_pressure_f64(device, _ptr(P_t), _ptr(G_t), int(Gs), _ptr(R_t), int(Rs), _ptr(phi_t),
                                     _ptr(theta_t), _ptr(area_t), int(npts), int(nbatch), int(lmax), int(mmax),
                                     float(rad_e), _lib.as_double_p(dfactor), _ptr(clm), _ptr(slm), _ptr(ws),
                                     ws.numel(), _stream_ptr())
    _lib.check(rc, 'mhb200_point_pressure_f64')
    return clm, slm


_angle_cache = []          # [(weakref(lon), lon._version, weakref(lat), lat._version, phi_t, theta_t)]


def _cached_point_angles(lon, lat, device):
    import weakref
    with _point_ws_lock:
        for rl, vl, ra, va, phi_t, theta_t in _angle_cache:
            if rl() is lon and ra() is lat and vl == lon._version and va == lat._version:
                return phi_t, theta_t
    phi_t = (_to_device(lon, device).reshape(-1) * (np.pi / 180.0)).contiguous()               # :111-113
    theta_t = ((90.0 - _to_device(lat, device).reshape(-1)) * (np.pi / 180.0)).contiguous()
    with _point_ws_lock:
        _angle_cache.insert(0, (weakref.ref(lon), lon._version, weakref.ref(lat), lat._version, phi_t, theta_t))
        del _angle_cache[4:]
    return phi_t, theta_t


def _flat(a):
    if isinstance(a, torch.Tensor):
        return a.reshape(-1)
    return np.asarray(a).flatten()


def _point_operand(a, device, npts, name):
    t = _to_device(_flat(a), device).contiguous()
    if t.numel() == 1:
        return t, 0
    if t.numel() != npts:
        raise ValueError('%s must have one value per point (or a single value)' % name)
    return t, 1


def _points_prepare(lon, lat, AREA, LMAX, MMAX, LOVE, device):
    if MMAX is None:                                                 # :107-108 (MMAX=0 stays 0)
        MMAX = np.copy(LMAX)
    if isinstance(lon, torch.Tensor) and isinstance(lat, torch.Tensor):
        # device-resident coordinates: np.radians(x) is the single multiply x*(pi/180), formed here
        # in fp64 on the device with the same constant; static point sets (the drivers' tiles) are
        # converted once: the cache holds the very tensor objects and their in-place version counters
        phi_t, theta_t = _cached_point_angles(lon, lat, device)
    else:
        phi_t = _to_device(np.radians(_flat(lon)), device)
        theta_t = _to_device(np.radians(90.0 - _flat(lat)), device)
    npts = phi_t.numel()
    factors = _compat.units(lmax=LMAX).spatial(*LOVE)                # :116-122
    rad_e = factors.rad_e / 100.0
    dfactor = 4.0 * np.pi * factors.mmwe / (1.0 + 2.0 * factors.l)
    area = AREA.flatten() if not isinstance(AREA, torch.Tensor) else AREA.reshape(-1)   # AttributeError for None, :126
    area_t, _ = _point_operand(area, device, npts, 'AREA')
    if area_t.numel() != npts:
        area_t = area_t.expand(npts).contiguous()
    return MMAX, npts, rad_e, dfactor, phi_t.contiguous(), theta_t.contiguous(), area_t


def gen_point_pressure(P, G, R, lon, lat, AREA=None, LMAX=60, MMAX=None, LOVE=None):
    r"""
    Stokes coefficients of point pressures with disc geometry
    (reference ``model_harmonics/gen_point_pressure.py:64-157``; same arguments).
    """
    device = require_device()
    MMAX, npts, rad_e, dfactor, phi_t, theta_t, area_t = _points_prepare(lon, lat, AREA, LMAX, MMAX, LOVE, device)
    P_t, _ = _point_operand(P, device, npts, 'P')
    if P_t.numel() != npts:
        P_t = P_t.expand(npts).contiguous()
    G_t, Gs = _point_operand(G, device, npts, 'G')
    R_t, Rs = _point_operand(R, device, npts, 'R')
    clm, slm = _points_core(device, P_t, G_t, Gs, R_t, Rs, phi_t, theta_t, area_t, npts, 1, int(LMAX), int(MMAX),
                            rad_e, dfactor)
    return _make_harmonics(clm[0].cpu().numpy(), slm[0].cpu().numpy(), LMAX, MMAX)


def point_pressure_batch(P, G, R, lon, lat, AREA=None, LMAX=60, MMAX=None, LOVE=None, as_numpy=True):
    """Batched form: ``P`` is (T, npts) over fixed points (reference ``OBP/ecco_llc_tile_harmonics.py:218-260``)."""
    device = require_device()
    MMAX, npts, rad_e, dfactor, phi_t, theta_t, area_t = _points_prepare(lon, lat, AREA, LMAX, MMAX, LOVE, device)
    P_t = _to_device(P, device).contiguous()
    if P_t.dim() != 2 or P_t.shape[1] != npts:
        raise ValueError('P must be (T, npts)')
    G_t, Gs = _point_operand(G, device, npts, 'G')
    R_t, Rs = _point_operand(R, device, npts, 'R')
    clm, slm = _points_core(device, P_t, G_t, Gs, R_t, Rs, phi_t, theta_t, area_t, npts, P_t.shape[0], int(LMAX),
                            int(MMAX), rad_e, dfactor)
    return (clm.cpu().numpy(), slm.cpu().numpy()) if as_numpy else (clm, slm)


def gen_point_pressure_tiles(tiles, LMAX=60, MMAX=None, LOVE=None):
    """
    Sum over spatial tiles of ``gen_point_pressure`` -- the accumulation loop of the LLC-tile driver
    (reference ``OBP/ecco_llc_tile_harmonics.py:235-260``: 13 calls whose harmonics are summed with
    ``obp_Ylms.add``).  ``tiles`` is a sequence of ``(P, G, R, lon, lat, AREA)`` (NumPy or device arrays,
    already reduced to the valid points of the tile).  The operator is linear in the points, so the tiles
    are concatenated on the device and transformed by ONE kernel pass; nothing goes back to the host
    between tiles.  ``P`` entries may carry a leading time axis (T, npts_k) for static-geometry batches:
    the result is then a pair of (T, LMAX+1, MMAX+1) arrays instead of a harmonics object.
    """
    device = require_device()
    dev = 'cuda:%d' % device
    cols = [[], [], [], [], [], []]
    batched = None
    for tile in tiles:
        P, G, R, lon, lat, AREA = tile
        P_t = _to_device(P, device)
        is_b = (P_t.dim() == 2)
        batched = is_b if batched is None else batched
        if is_b != batched:
            raise ValueError('all tiles must be batched (T, npts) or all unbatched')
        n = P_t.shape[-1] if is_b else P_t.numel()
        cols[0].append(P_t if is_b else P_t.reshape(-1))
        for k, a in ((1, G), (2, R), (3, lon), (4, lat), (5, AREA)):
            t = _to_device(_flat(a), device).reshape(-1)
            cols[k].append(t.expand(n) if t.numel() == 1 else t)
            if cols[k][-1].numel() != n:
                raise ValueError('tile operands must have one value per point (or a single value)')
    if not cols[0]:
        raise ValueError('no tiles')
    P_all = torch.cat(cols[0], dim=-1)
    G_all, R_all, lon_all, lat_all, A_all = (torch.cat(c) for c in cols[1:])
    if batched:
        return point_pressure_batch(P_all, G_all, R_all, lon_all, lat_all, AREA=A_all, LMAX=LMAX, MMAX=MMAX,
                                    LOVE=LOVE)
    return gen_point_pressure(P_all, G_all, R_all, lon_all, lat_all, AREA=A_all, LMAX=LMAX, MMAX=MMAX, LOVE=LOVE)
