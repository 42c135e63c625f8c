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
        mma = (5120.0 if lb == mb else 8192.0) / (1 << fold)
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
    sets = np.zeros(4 * n, dtype=np.int32)
    dummy = np.zeros(1, dtype=np.int32)
    fold, nchunks, kcf = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    nf = _lib.lib().mhb200_fold_probe(n, _lib.as_double_p(np.ascontiguousarray(phi)), nth, sets.ctypes.data_as(INT_P),
                                      dummy.ctypes.data_as(INT_P), ctypes.byref(fold), ctypes.byref(nchunks),
                                      ctypes.byref(kcf))
    assert nf > 0
    return phi, sets[:4 * nf].reshape(nf, 4), fold.value, nchunks.value, kcf.value


def _folded_sums(phi, sets, fold, B, mmax):
    """Fourier sums over the contraction columns exactly as the kernel forms them."""
    dc, ds = np.zeros(mmax + 1), np.zeros(mmax + 1)
    for p, q, p2, q2 in sets:
        x1 = B[p]
        x2 = B[q] if fold >= 1 else 0.0
        x3 = B[p2] if fold >= 2 else 0.0
        x4 = B[q2] if fold >= 2 else 0.0
        w = 1.0
        if fold >= 1 and q == p:
            w = 0.5
        if fold >= 2 and p2 == p:
            w = 0.5
        s12, d12, s34, d34 = x1 + x2, x1 - x2, x3 + x4, x3 - x4
        for m in range(mmax + 1):
            if fold == 0:
                bc, bs = x1, x1
            elif fold == 1:
                bc, bs = s12, d12
            else:
                bc, bs = (s12 + s34, d12 - d34) if m % 2 == 0 else (s12 - s34, d12 + d34)
            dc[m] += w * np.cos(m * phi[p]) * bc
            ds[m] += w * np.sin(m * phi[p]) * bs
    return dc, ds


@pytest.mark.parametrize('lon,level', [(np.arange(1440) * 0.25, 2), (np.arange(360) * 1.0, 2),
                                       (-180.0 + np.arange(120) * 3.0, 2), (np.arange(-179.5, 180.0, 1.0), 2),
                                       (np.arange(7) * (360.0 / 7), 1), (np.arange(90) * 4.0 - 180.0, 2),
                                       (np.arange(18) * 20.0, 2), (np.arange(10) * 36.0, 2)])
@pytest.mark.parametrize('nth', [128, 256])
def test_regular_grids_fold_and_reproduce_the_fourier_sums(lon, level, nth):
    phi, sets, fold, nchunks, kcf = _fold(lon, nth)
    n = len(phi)
    assert fold == level
    # every grid column appears in exactly one contraction column
    seen = np.zeros(n, dtype=int)
    for row in sets:
        for c in set(int(v) for v in row[:1 << fold]):
            seen[c] += 1
    assert np.all(seen == 1)
    # the folded sums equal the direct ones
    rng = np.random.default_rng(n)
    B = rng.standard_normal(n)
    mmax = 40
    dc, ds = _folded_sums(phi, sets, fold, B, mmax)
    m = np.arange(mmax + 1)[:, None]
    want_c, want_s = np.cos(m * phi[None, :]) @ B, np.sin(m * phi[None, :]) @ B
    scale = np.abs(want_c).max()
    assert np.abs(dc - want_c).max() <= 1e-12 * scale and np.abs(ds - want_s).max() <= 1e-12 * scale
    # chunking holds all contraction columns and fits the CTA shape
    assert kcf % 8 == 0 and kcf <= (nth // 2) >> fold and nchunks * kcf >= len(sets)


def test_asymmetric_grid_does_not_fold():
    phi, sets, fold, nchunks, kcf = _fold(np.arange(90) * 4.0 + 0.7, 256)
    assert fold == 0 and np.all(sets[:, 1:] == -1) and np.array_equal(sets[:, 0], np.arange(90))
    assert kcf <= 128 and nchunks * kcf >= 90


def test_harmonics_ascii_round_trip(tmp_path):
    """SURVEY 8f-4, the part that needs neither netCDF4 nor h5py: the text format of
    harmonics.to_ascii / from_ascii (one 'l m clm slm [time]' line per coefficient, order-major)."""
    from model_harmonics_b200 import _compat
    if _compat.HAVE_GRAVITY_TOOLKIT:
        pytest.skip('the real gravity_toolkit container is in use')
    rng = np.random.default_rng(5)
    L, M = 12, 7
    Y = _compat.harmonics(lmax=L, mmax=M)
    Y.clm = np.tril(rng.standard_normal((L + 1, L + 1)))[:, :M + 1] * 1e-9
    Y.slm = np.tril(rng.standard_normal((L + 1, L + 1)), -1)[:, :M + 1] * 1e-9
    Y.time = 2003.4567
    path = tmp_path / 'ERA5_CLM_L12M7_016.txt'
    Y.to_file(path, format='ascii')
    lines = path.read_text().splitlines()
    assert len(lines) == sum(L + 1 - m for m in range(M + 1))
    assert lines[0] == '{0:5d} {1:5d} {2:+21.12e} {3:+21.12e} {4:10.4f}'.format(0, 0, Y.clm[0, 0], 0.0, 2003.4567)
    Z = _compat.harmonics().from_file(path, format='ascii')
    assert (int(Z.lmax), int(Z.mmax)) == (L, M) and Z.time == 2003.4567
    # 13 significant digits survive the text form
    assert np.allclose(Z.clm, Y.clm, rtol=1e-12, atol=0.0) and np.allclose(Z.slm, Y.slm, rtol=1e-12, atol=0.0)
    with pytest.raises(ValueError):
        Y.to_file(path, format='netCDF4')
    # without a date column
    Y.to_ascii(path, date=False)
    assert len(path.read_text().splitlines()[0].split()) == 4
    W = _compat.harmonics().from_ascii(path, date=False)
    assert np.allclose(W.clm, Y.clm, rtol=1e-12, atol=0.0)


def test_moment_schedule_probe_partitions_units_and_slots():
    """Static schedule of the contracted-moment atmosphere kernel (csrc/stokes_moments.cu)."""
    import ctypes
    from model_harmonics_b200 import _lib
    L = _lib.lib()
    for nb, nrings, nchunks, max_ctas in [(1, 721, 6, 296), (8, 721, 6, 296), (1, 46, 1, 296), (3, 7, 2, 5),
                                          (2, 90, 6, 296), (1, 1, 1, 296)]:
        cu = (ctypes.c_int * (max_ctas + 1))()
        cs = (ctypes.c_int * max_ctas)()
        fs = (ctypes.c_int * (nb + 1))()
        ns = ctypes.c_int(0)
        ncta = L.mhb200_moment_schedule_probe(nb, nrings, nchunks, max_ctas, cu, cs, fs, max_ctas + 1,
                                              ctypes.byref(ns))
        units = nb * nrings * nchunks
        assert ncta == min(max_ctas, units)
        assert cu[0] == 0 and cu[ncta] == units
        spans = [cu[b + 1] - cu[b] for b in range(ncta)]
        assert min(spans) >= 1 and max(spans) - min(spans) <= 1          # equal split
        # slots: one per (CTA, ring) pair; consecutive and grouped by field
        slot = 0
        rings_of_field = [set() for _ in range(nb)]
        for b in range(ncta):
            assert cs[b] == slot
            rings = sorted(set(u // nchunks for u in range(cu[b], cu[b + 1])))
            slot += len(rings)
            for r in rings:
                rings_of_field[r // nrings].add(r % nrings)
        assert ns.value == slot
        assert fs[0] == 0 and fs[nb] == slot
        assert all(fs[t] <= fs[t + 1] for t in range(nb))
        assert all(len(s) == nrings for s in rings_of_field)
        assert slot <= nb * nrings + ncta


def test_standard_plm_table_matches_the_oracle_table_bitwise():
    from model_harmonics_b200 import _compat
    from oracle.gravtk_standin import plm_holmes
    x = np.cos(np.radians(90.0 - np.linspace(90.0, -90.0, 19)))
    mine = _compat.plm_table(30, 17, x)
    ref, _ = plm_holmes(30, x)
    assert np.array_equal(mine, ref[:, :18, :])


@pytest.mark.parametrize('lon,tma', [
    (np.arange(1440) * 0.25, True), (np.arange(360) * 1.0, True), (-180.0 + np.arange(120) * 3.0, True),
    (np.arange(-179.5, 180.0, 1.0), True), (np.arange(8) * 45.0, True), (np.arange(260) * (360.0 / 260), True),
    # rotated by three steps: both symmetries hold, but the fold's base columns are not one contiguous run
    (3 * (360.0 / 516) + np.arange(516) * (360.0 / 516), False)])
def test_moment_layout_probe_covers_every_column_once_with_aligned_runs(lon, tma):
    """Chunk layout of the contracted-moment atmosphere kernel (csrc/stokes_moments.cu, moment_layout):
    every grid column is owned by exactly one active thread column (duplicates of self-paired columns are
    dropped), each fold sub-block of a chunk is one longitude run starting at box_start, and the kernel's
    Fourier sums over the kept columns -- weight 1 everywhere -- reproduce the direct sums."""
    phi, sets, fold, _, _ = _fold(lon, 256)
    assert fold == 2
    n, nf = len(phi), len(sets)
    nchunks = (nf + 63) // 64
    cg = np.zeros(nchunks * 256, dtype=np.int32)
    bs = np.zeros(nchunks * 4, dtype=np.int32)
    dr = np.zeros(4, dtype=np.int32)
    rc = _lib.lib().mhb200_moment_layout_probe(nf, np.ascontiguousarray(sets.reshape(-1)).ctypes.data_as(INT_P),
                                               cg.ctypes.data_as(INT_P), bs.ctypes.data_as(INT_P),
                                               dr.ctypes.data_as(INT_P), nchunks)
    assert rc == (1 if tma else 0)
    cg = cg.reshape(nchunks, 8, 4, 8)              # [chunk][warp][sub][lane % 8]
    owned = cg[cg >= 0]
    assert len(owned) == n and len(set(owned.tolist())) == n          # every grid column exactly once
    rng = np.random.default_rng(n)
    B = rng.standard_normal(n)
    mmax = 30
    dc, ds = np.zeros(mmax + 1), np.zeros(mmax + 1)
    for c in range(nchunks):
        for w in range(8):
            for jo in range(8):
                j = 64 * c + 8 * w + jo
                if j >= nf:
                    assert np.all(cg[c, w, :, jo] < 0)
                    continue
                x = [B[col] if col >= 0 else 0.0 for col in cg[c, w, :, jo]]
                s12, d12, s34, d34 = x[0] + x[1], x[0] - x[1], x[2] + x[3], x[2] - x[3]
                p = sets[j, 0]
                assert cg[c, w, 0, jo] == p                             # sub-block 0 is never dropped
                for m in range(mmax + 1):
                    bc, bsn = (s12 + s34, d12 - d34) if m % 2 == 0 else (s12 - s34, d12 + d34)
                    dc[m] += np.cos(m * phi[p]) * bc
                    ds[m] += np.sin(m * phi[p]) * bsn
                if tma:
                    for sb in range(4):
                        col = cg[c, w, sb, jo]
                        if col >= 0:
                            jj = 8 * w + jo
                            pos = jj if dr[sb] > 0 else 63 - jj
                            assert col == bs[4 * c + sb] + pos
    m = np.arange(mmax + 1)[:, None]
    want_c, want_s = np.cos(m * phi[None, :]) @ B, np.sin(m * phi[None, :]) @ B
    scale = np.abs(want_c).max()
    assert np.abs(dc - want_c).max() <= 1e-12 * scale and np.abs(ds - want_s).max() <= 1e-12 * scale
