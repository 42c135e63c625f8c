"""
ctypes binding of ``libmhb200.so`` (C ABI declared in ``include/mhb200.h``).

There is no CPU fallback: if the library is missing or no sm_100 device is present the
operators raise.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('MHB200_LIB', os.path.join(_HERE, 'libmhb200.so'))

OK = 0
METHOD_BC05, METHOD_SW02 = 0, 1
F64, F32 = 0, 1

_lib = None

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int64_p = ctypes.POINTER(ctypes.c_int64)
_vp = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/mhb200.h one to one
PROTOTYPES = {
    'mhb200_version': (ctypes.c_int, []),
    'mhb200_device_check': (ctypes.c_int, [ctypes.c_int]),
    'mhb200_last_error': (ctypes.c_char_p, []),
    'mhb200_launch_count': (ctypes.c_uint64, []),
    'mhb200_profile_enable': (None, [ctypes.c_int]),
    'mhb200_profile_mean_ms': (ctypes.c_double, [ctypes.POINTER(ctypes.c_int)]),
    'mhb200_plan_create': (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_int, _c_double_p, _c_double_p, _c_double_p]),
    'mhb200_plan_destroy': (ctypes.c_int, [_vp]),
    'mhb200_grid_workspace_bytes': (ctypes.c_size_t, [_vp, ctypes.c_int, ctypes.c_int]),
    'mhb200_pressure_stokes_f64': (ctypes.c_int, [_vp, _vp, _c_int64_p, _vp, _c_int64_p, _vp, _c_int64_p,
                                                  ctypes.c_int, ctypes.c_double, _c_double_p, _vp, _vp, _vp,
                                                  _vp, ctypes.c_size_t, _vp]),
    'mhb200_atmosphere_stokes': (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, ctypes.c_int64, _c_double_p, _vp,
                                                _c_int64_p, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_double, _c_double_p, _vp, _vp, _vp, _vp,
                                                ctypes.c_size_t, _vp]),
    'mhb200_host_workspace_bytes': (ctypes.c_size_t, [_vp, ctypes.c_int]),
    'mhb200_atmosphere_stokes_host': (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, _vp, _c_int64_p, _vp, _vp,
                                                     ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                                     _c_double_p, _vp, _c_double_p, _c_double_p, _vp,
                                                     ctypes.c_size_t, _vp]),
    'mhb200_atmosphere_stokes_host_geoid': (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, _c_double_p, _vp, _vp,
                                                           ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                                           _c_double_p, _vp, _c_double_p, _c_double_p, _vp,
                                                           ctypes.c_size_t, _vp]),
    'mhb200_geopotential_levels_f32': (ctypes.c_int, [ctypes.c_int, _vp, _vp, _vp, _vp, _c_double_p, _c_double_p,
                                                      ctypes.c_int, _c_double_p, _c_double_p, ctypes.c_int, ctypes.c_int,
                                                      ctypes.c_int64, ctypes.c_double, ctypes.c_double, ctypes.c_float,
                                                      _vp, _vp, _vp]),
    'mhb200_schedule_probe': (ctypes.c_int, [ctypes.c_int] * 11 + [ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                                                              ctypes.POINTER(ctypes.c_int), ctypes.c_int,
                                                              ctypes.POINTER(ctypes.c_int)]),
    'mhb200_fold_probe': (ctypes.c_int, [ctypes.c_int, _c_double_p, ctypes.c_int] + [ctypes.POINTER(ctypes.c_int)] * 5),
    'mhb200_moment_schedule_probe': (ctypes.c_int, [ctypes.c_int] * 4 + [ctypes.POINTER(ctypes.c_int)] * 3 +
                                     [ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    'mhb200_moment_layout_probe': (ctypes.c_int, [ctypes.c_int] + [ctypes.POINTER(ctypes.c_int)] * 4 + [ctypes.c_int]),
    'mhb200_sincos_probe': (ctypes.c_int, [ctypes.c_int, _vp, ctypes.c_int64, _vp, _vp, _vp]),
    'mhb200_points_workspace_bytes': (ctypes.c_size_t, [ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    'mhb200_point_pressure_f64': (ctypes.c_int, [ctypes.c_int, _vp, _vp, ctypes.c_int64, _vp, ctypes.c_int64,
                                                 _vp, _vp, _vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_int, ctypes.c_double, _c_double_p, _vp, _vp, _vp,
                                                 ctypes.c_size_t, _vp]),
}


def lib():
    """Loads the shared library once; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                'model_harmonics_b200: %s not found -- run build.sh (or __graft_entry__.build()); '
                'there is no CPU fallback' % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what=''):
    if rc != OK:
        msg = lib().mhb200_last_error().decode('utf-8', 'replace')
        raise RuntimeError('%s failed with status %d: %s' % (what or 'mhb200 call', rc, msg))


def as_double_p(a):
    return a.ctypes.data_as(_c_double_p)


def int64x(*vals):
    return (ctypes.c_int64 * len(vals))(*[int(v) for v in vals])
