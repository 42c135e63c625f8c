"""
Geometry plans: device-resident tables that depend only on (lon, lat, LMAX, MMAX, weights),
cached so that a driver's per-time-step calls (reference
``reanalysis/reanalysis_pressure_harmonics.py:472-503``) pay for them once.
"""
import ctypes
import hashlib
import os
import threading
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from . import _compat

_MAX_PLANS = int(os.environ.get('MHB200_PLAN_CACHE', '8'))
_cache = OrderedDict()
_cache_lock = threading.Lock()


def require_device():
    if not torch.cuda.is_available():
        raise RuntimeError('model_harmonics_b200 needs a CUDA device (B200, sm_100a); '
                           'there is no CPU fallback')
    return torch.cuda.current_device()


def _digest(*arrays):
    h = hashlib.sha1()
    for a in arrays:
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return h.hexdigest()


class GridPlan(object):
    """Owns one ``mhb200_plan`` plus workspaces and per-plan device tables."""

    def __init__(self, device, phi, costh, int_fact, lmax, mmax):
        L = _lib.lib()
        self.device = int(device)
        self.nlat, self.nlon = len(costh), len(phi)
        self.lmax, self.mmax = int(lmax), int(mmax)
        phi = np.ascontiguousarray(phi, dtype=np.float64)
        costh = np.ascontiguousarray(costh, dtype=np.float64)
        int_fact = np.ascontiguousarray(int_fact, dtype=np.float64)
        handle = ctypes.c_void_p()
        _lib.check(L.mhb200_plan_create(ctypes.byref(handle), self.device, self.nlat, self.nlon,
                                        self.lmax, self.mmax, _lib.as_double_p(phi),
                                        _lib.as_double_p(costh), _lib.as_double_p(int_fact)),
                   'mhb200_plan_create')
        self.handle = handle
        self.costh = costh
        self._workspaces = {}     # CUDA stream handle -> workspace tensor
        self.tables = {}          # per-plan device tensors (atmosphere ring tables, PLM tables)
        self.lock = threading.Lock()

    def _workspace_for_stream(self, need):
        """
        One workspace per (plan, CUDA stream): the kernels of a call keep partial sums and staged
        operands there, so calls enqueued on different streams (two threads, or one thread switching
        streams) must not share it (re-entrant per (device, stream)).  Same-stream calls
        are ordered by the stream and reuse the buffer.
        """
        key = int(torch.cuda.current_stream().cuda_stream)
        with self.lock:
            ws = self._workspaces.get(key)
            if ws is None or ws.numel() < need:
                ws = torch.empty(need, dtype=torch.uint8, device='cuda:%d' % self.device)
                self._workspaces[key] = ws
        return ws

    def workspace(self, nbatch, nlev=0):
        need = _lib.lib().mhb200_grid_workspace_bytes(self.handle, int(nbatch), int(nlev))
        if need == 0:
            raise ValueError('bad batch size')
        return self._workspace_for_stream(need)

    def host_workspace(self, nbands):
        return self._workspace_for_stream(_lib.lib().mhb200_host_workspace_bytes(self.handle, int(nbands)))

    def close(self):
        if getattr(self, 'handle', None) is not None and self.handle.value:
            try:
                _lib.lib().mhb200_plan_destroy(self.handle)
            except Exception:
                pass
            self.handle = None

    def __del__(self):
        self.close()


def get_plan(phi, costh, int_fact, lmax, mmax, device=None):
    device = require_device() if device is None else device
    key = (device, int(lmax), int(mmax), _digest(phi, costh, int_fact))
    with _cache_lock:
        plan = _cache.get(key)
        if plan is not None:
            _cache.move_to_end(key)
            return plan
    plan = GridPlan(device, phi, costh, int_fact, lmax, mmax)
    with _cache_lock:
        _cache[key] = plan
        while len(_cache) > _MAX_PLANS:
            _cache.popitem(last=False)
    return plan


def clear_plans():
    with _cache_lock:
        _cache.clear()


def plm_table_for(plan, PLM, lmax, mmax):
    """
    Decides how to honour a caller-supplied ``PLM`` (reference ``gen_pressure_stokes.py:180-188``:
    the reference always contracts with the table it is given).
    ``MHB200_PLM_MODE``: ``auto`` (default) compares the WHOLE table ``PLM[:L+1, :M+1, :]`` with the
    standard Holmes-Featherstone values for ``lat`` (host table cached per plan) and regenerates it on
    the fly only when every entry agrees to 1e-12 of the table's scale -- otherwise the caller's table is
    uploaded and used; ``table`` always uses the table; ``recompute`` never does.
    The verdict is cached per (buffer address, shape, strides, strided content sample), so a driver that
    passes the same array every time step pays for the comparison once.
    Returns None (on-the-fly recursion) or a (nlat, lmax+1, mmax+1) device tensor.
    """
    if PLM is None:
        return None
    mode = os.environ.get('MHB200_PLM_MODE', 'auto').lower()
    if mode == 'recompute':
        return None
    arr = np.asarray(PLM)
    flat = arr.reshape(-1)
    sample = np.ascontiguousarray(flat[::max(1, flat.size // 8192)], dtype=np.float64)
    key = ('plm', arr.__array_interface__['data'][0], arr.shape, arr.strides, str(arr.dtype),
           hashlib.sha1(sample.tobytes()).hexdigest(), lmax, mmax, mode)
    with plan.lock:
        if key in plan.tables:
            return plan.tables[key]
    use_table = (mode == 'table')
    if not use_table:
        skey = ('plm_standard', lmax, mmax)
        with plan.lock:
            std = plan.tables.get(skey)
        if std is None:
            std = _compat.plm_table(lmax, mmax, plan.costh)
            with plan.lock:
                plan.tables[skey] = std
        given = arr[:lmax + 1, :mmax + 1, :]
        if given.shape != std.shape:
            raise ValueError('PLM must be at least (LMAX+1, MMAX+1, nlat); got %s' % (arr.shape,))
        scale = max(float(np.abs(std).max()), 1.0)
        diff = np.abs(given - std)
        # NaNs in the caller's table count as "different": the table is then used as given
        use_table = not bool(np.all(diff <= 1e-12 * scale))
    table = None
    if use_table:
        host = np.ascontiguousarray(np.transpose(np.asarray(PLM, dtype=np.float64)[:lmax + 1, :mmax + 1, :],
                                                 (2, 0, 1)))
        table = torch.from_numpy(host).to('cuda:%d' % plan.device)
    with plan.lock:
        plan.tables[key] = table
    return table
