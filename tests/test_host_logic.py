"""
CPU tests of the library's pure host logic through the C-ABI probes (no device needed): the static
equal-cost schedule and the longitude mirror pairing / chunking.
"""
import ctypes

import numpy as np
import pytest

from model_harmonics_b200 import _lib

INT_P = ctypes.POINTER(ctypes.c_int)


def _schedule(nlat, nchunks, nlb, nmb, nbatch, max_ctas, mode=1, nlev=37, fold=1, h0=0, h1=None):
    h1 = nlat if h1 is None else h1
    seg = np.zeros(5 * 20000, dtype=np.int32)
    cta = np.zeros(max_ctas + 2, dtype=np.int32)
    ncta = ctypes.c_int(0)
    n = _lib.lib().mhb200_schedule_probe(nlat, nchunks, nlb, nmb, nbatch, max_ctas, mode, nlev, fold, h0, h1,
                                         seg.ctypes.data_as(INT_P), 20000, cta.ctypes.data_as(INT_P), max_ctas + 2,
                                         ctypes.byref(ncta))
    assert n > 0
    return seg[:5 * n].reshape(n, 5), cta[:ncta.value + 1], ncta.value


@pytest.mark.parametrize('shape', [
    dict(nlat=721, nchunks=23, nlb=1, nmb=1, nbatch=1, max_ctas=296),       # C3, 128-thread shape
    dict(nlat=721, nchunks=23, nlb=1, nmb=1, nbatch=8, max_ctas=296),
    dict(nlat=181, nchunks=2, nlb=1, nmb=1, nbatch=12, max_ctas=148, mode=0),
    dict(nlat=721, nchunks=12, nlb=6, nmb=6, nbatch=1, max_ctas=148, mode=0),  # C4: 21 tiles, mixed costs
    dict(nlat=7, nchunks=1, nlb=1, nmb=1, nbatch=1, max_ctas=148),            # fewer units than CTAs
    dict(nlat=91, nchunks=3, nlb=2, nmb=1, nbatch=3, max_ctas=148, mode=0),   # MMAX < LMAX blocks
    dict(nlat=721, nchunks=23, nlb=1, nmb=1, nbatch=1, max_ctas=296, h0=90, h1=180),   # one latitude band
])
def test_schedule_covers_every_unit_once_and_balances(shape):
    seg, cta, ncta = _schedule(**shape)
    nlat, nchunks, nlb, nmb, nbatch = (shape[k] for k in ('nlat', 'nchunks', 'nlb', 'nmb', 'nbatch'))
    h0, h1 = shape.get('h0', 0), shape.get('h1', nlat)
    tiles = [(lb, mb) for lb in range(nlb) for mb in range(min(lb, nmb - 1) + 1)]
    # segments are ordered by (field, tile, unit) and tile the unit range of each (field, tile) exactly
    i = 0
    for t in range(nbatch):
        for lb, mb in tiles:
            u = h0 * nchunks
            while u < h1 * nchunks:
                assert tuple(seg[i][:3]) == (t, lb, mb) and seg[i][3] == u and seg[i][4] > u
                u = seg[i][4]
                i += 1
            assert u == h1 * nchunks
    assert i == len(seg)
    # CTA lists are contiguous and cover all segments
    assert cta[0] == 0 and cta[-1] == len(seg) and np.all(np.diff(cta) >= 0)
    # equal-cost split: per-CTA cost within one unit of the mean (cost model of ring_cost)
    mode, nlev, fold = shape.get('mode', 1), shape.get('nlev', 37), shape.get('fold', 1)
    def cost(lb, mb):
        mma = (4608.0 if lb == mb else 8192.0) * (0.5 if fold else 1.0)
        return mma + (90.0 if mode == 0 else nlev * 136.0)
    per = np.array([sum((s[4] - s[3]) * cost(s[1], s[2]) for s in seg[cta[c]:cta[c + 1]]) for c in range(ncta)])
    if per.sum() > 0 and ncta > 1:
        biggest = max(cost(lb, mb) for lb, mb in tiles)
        assert per.max() - per.mean() <= 1.01 * biggest
        assert np.count_nonzero(per) >= min(ncta, len(seg)) - 1


def _fold(lon_deg, nth, wrap=True):
    phi = np.radians(np.asarray(lon_deg, dtype=np.float64))
    if wrap:
        phi = np.where(phi < 0, phi + 2.0 * np.pi, phi)
    n = len(phi)
    colp = np.zeros(n, dtype=np.int32)
    colq = np.zeros(n, dtype=np.int32)
    fold, nchunks, kcf = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    nf = _lib.lib().mhb200_fold_probe(n, _lib.as_double_p(np.ascontiguousarray(phi)), nth, colp.ctypes.data_as(INT_P),
                                      colq.ctypes.data_as(INT_P), ctypes.byref(fold), ctypes.byref(nchunks),
                                      ctypes.byref(kcf))
    assert nf > 0
    return phi, colp[:nf], colq[:nf], fold.value, nchunks.value, kcf.value


@pytest.mark.parametrize('lon', [np.arange(1440) * 0.25, np.arange(360) * 1.0, -180.0 + np.arange(120) * 3.0,
                                 np.arange(-179.5, 180.0, 1.0), np.arange(7) * (360.0 / 7), np.arange(90) * 4.0 - 180.0])
@pytest.mark.parametrize('nth', [128, 256])
def test_regular_grids_fold_into_mirror_pairs(lon, nth):
    phi, colp, colq, fold, nchunks, kcf = _fold(lon, nth)
    n = len(phi)
    assert fold == 1
    # every grid column appears exactly once (self-mirrored ones as their own partner)
    seen = np.zeros(n, dtype=int)
    for p, q in zip(colp, colq):
        seen[p] += 1
        if q != p:
            seen[q] += 1
        # mirror relation: phi_p + phi_q is a multiple of 2 pi to a few ulp
        d = (phi[p] + phi[q]) / (2.0 * np.pi)
        assert abs(d - round(d)) < 1e-15
    assert np.all(seen == 1)
    # chunking holds all contraction columns with little padding and fits the CTA shape
    assert kcf % 8 == 0 and kcf <= nth // 4 and nchunks * kcf >= len(colp)
    assert nchunks * kcf - len(colp) < max(0.08 * len(colp), 8 * nchunks)


def test_asymmetric_grid_does_not_fold():
    phi, colp, colq, fold, nchunks, kcf = _fold(np.arange(90) * 4.0 + 0.7, 256)
    assert fold == 0 and np.all(colq == -1) and np.array_equal(colp, np.arange(90))
    assert kcf <= 128 and nchunks * kcf >= 90
