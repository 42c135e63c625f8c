"""
Device-side forms of the small per-time-step operations the reference drivers wrap around the
operators, so that a multi-month job stays on the GPU between the file read and the
final gather:

* pressure anomaly            ``reanalysis_pressure_harmonics.py:474``  ``P = pressure[t] - mean``
* ocean redistribution        ``reanalysis_pressure_harmonics.py:476-490`` and, per level,
                              ``reanalysis_atmospheric_harmonics.py:361-379``: replace the values at the
                              ocean points by their area-weighted mean
* mean-harmonics removal      ``reanalysis_atmospheric_harmonics.py:396``  ``Ylms.subtract(mean_Ylms)``
* multi-month mean            ``reanalysis_mean_harmonics.py:381``  ``from_list(...).mean()``

These are elementwise / small-reduction steps off the hot path; they are expressed with torch tensor
ops on the caller's device tensors (no host round trip) and feed the batched operators directly.
"""
import numpy as np
import torch

__all__ = ['pressure_anomaly', 'redistribute_ocean', 'subtract_mean_harmonics', 'mean_harmonics']


def _dev(a, like=None, dtype=torch.float64):
    if isinstance(a, torch.Tensor):
        return a
    t = torch.from_numpy(np.ascontiguousarray(a))
    if t.dtype.is_floating_point and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(like.device) if like is not None else t.cuda()


def pressure_anomaly(pressure, mean_pressure):
    """``pressure`` (..., nlat, nlon) minus the (nlat, nlon) mean field."""
    p = _dev(pressure)
    return p - _dev(mean_pressure, like=p, dtype=p.dtype)


def redistribute_ocean(field, area, ocean_mask):
    """
    Replaces ``field[..., ii, jj]`` at the ocean points by ``sum(field*area)/sum(area)`` over those
    points, separately for every leading index (time step, level).  ``ocean_mask`` is the boolean
    (nlat, nlon) form of the drivers' ``(ii, jj)`` index lists.  Returns a new tensor.
    """
    f = _dev(field)
    a = _dev(area, like=f, dtype=f.dtype)
    m = _dev(np.asarray(ocean_mask, dtype=bool) if not isinstance(ocean_mask, torch.Tensor) else ocean_mask, like=f)
    am = torch.where(m, a, torch.zeros_like(a))
    total = am.sum()
    newtons = (f * am).sum(dim=(-2, -1), keepdim=True)
    return torch.where(m, newtons / total, f)


def subtract_mean_harmonics(clm, slm, mean_clm, mean_slm):
    """(clm, slm) (..., L+1, M+1) minus mean harmonics, truncated to the common degree and order."""
    c, s = _dev(clm), _dev(slm)
    mc, ms = _dev(mean_clm, like=c), _dev(mean_slm, like=c)
    l1 = min(c.shape[-2], mc.shape[-2])
    m1 = min(c.shape[-1], mc.shape[-1])
    c, s = c.clone(), s.clone()
    c[..., :l1, :m1] -= mc[:l1, :m1]
    s[..., :l1, :m1] -= ms[:l1, :m1]
    return c, s


def mean_harmonics(clm, slm):
    """Mean over the leading (time) axis of batched coefficients."""
    return _dev(clm).mean(dim=0), _dev(slm).mean(dim=0)
