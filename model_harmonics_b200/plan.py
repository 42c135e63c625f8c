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
        self._workspace = None
        self.tables = {}          # per-plan device tensors (atmosphere ring tables, PLM tables)
        self.lock = threading.Lock()

    def workspace(self, nbatch, nlev=0):
        need = _lib.lib().mhb200_grid_workspace_bytes(self.handle, int(nbatch), int(nlev))
        if need == 0:
            raise ValueError('bad batch size')
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = torch.empty(need, dtype=torch.uint8, device='cuda:%d' % self.device)
        return self._workspace

    def host_workspace(self, nbands):
        need = _lib.lib().mhb200_host_workspace_bytes(self.handle, int(nbands))
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = torch.empty(need, dtype=torch.uint8, device='cuda:%d' % self.device)
        return self._workspace

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
    Decides how to honour a caller-supplied ``PLM`` (reference ``gen_pressure_stokes.py:180-188``).
    ``MHB200_PLM_MODE``: ``auto`` (default) recomputes on the fly when the table matches the
    standard Holmes-Featherstone values for ``lat`` at sampled entries, otherwise uploads and
    uses the caller's table; ``table`` always uses the table; ``recompute`` never does.
    Returns None (on-the-fly recursion) or a (nlat, lmax+1, mmax+1) device tensor.
    """
    if PLM is None:
        return None
    mode = os.environ.get('MHB200_PLM_MODE', 'auto').lower()
    if mode == 'recompute':
        return None
    # content fingerprint (strided sample), not object identity: addresses get reused
    flat = np.asarray(PLM).reshape(-1)
    sample = np.ascontiguousarray(flat[::max(1, flat.size // 8192)], dtype=np.float64)
    key = ('plm', hashlib.sha1(sample.tobytes()).hexdigest(), tuple(np.shape(PLM)), lmax, mmax)
    with plan.lock:
        if key in plan.tables:
            return plan.tables[key]
    use_table = (mode == 'table')
    if not use_table:
        samples = sorted(set([(0, 0), (min(1, lmax), 0), (lmax, 0), (lmax, min(mmax, lmax)),
                              (lmax, min(mmax, lmax // 2)), (max(lmax // 2, min(mmax, 1)), min(mmax, 1))]))
        samples = [(l, m) for (l, m) in samples if m <= l]
        hs = sorted(set(int(v) for v in np.linspace(0, plan.nlat - 1, 7)))
        ref = _compat.plm_columns(lmax, mmax, plan.costh[hs], samples)
        got = np.array([[float(PLM[l, m, h]) for h in hs] for (l, m) in samples])
        scale = np.maximum(np.abs(ref).max(), 1.0)
        use_table = bool(np.abs(ref - got).max() > 1e-12 * scale)
    table = None
    if use_table:
        host = np.ascontiguousarray(np.transpose(np.asarray(PLM, dtype=np.float64)[:lmax + 1, :mmax + 1, :],
                                                 (2, 0, 1)))
        table = torch.from_numpy(host).to('cuda:%d' % plan.device)
    with plan.lock:
        plan.tables[key] = table
    return table
