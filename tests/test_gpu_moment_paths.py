"""
GPU tests of the contracted-moment kernels (csrc/stokes_moments.cu) and of their guard: the binomial
expansion is valid for |r / r0 - 1| <= 0.8 / n_max; inputs outside that range must come back exact through
the direct kernels (the reference's np.power is exact for any radius, gen_atmosphere_stokes.py:265,
gen_pressure_stokes.py:200).  Also: flipped (negative-stride) operands and whole-table PLM validation.
"""
import numpy as np
import pytest

from parity import TOL, rel_to_max, assert_structure

pytestmark = pytest.mark.gpu


def _mh():
    import model_harmonics_b200 as mh
    return mh


def _launches():
    from model_harmonics_b200 import _lib
    return _lib.lib().mhb200_launch_count()


@pytest.mark.parametrize('method', ['BC05', 'SW02'])
@pytest.mark.parametrize('top_km', [48.0, 80.0, 150.0])
def test_atmosphere_heights_inside_and_outside_the_guard(method, top_km):
    """48 and 80 km stay on the moment path; 150 km trips the guard and is recomputed by the direct kernels."""
    from oracle import stokes_oracle
    from model_harmonics_b200 import testing as syn
    mh = _mh()
    L = 60
    lon, lat = syn.grid_axes(46, 90)
    GPH, dP, geoid = syn.atmosphere_cube(37, 46, 90, seed=31)
    GPH = GPH * (top_km / 48.0)
    kw = dict(LMAX=L, ELLIPSOID='WGS84', GEOID=geoid, LOVE=syn.love_numbers(L), METHOD=method)
    Y = mh.gen_atmosphere_stokes(GPH, dP, lon, lat, **kw)
    O = stokes_oracle.gen_atmosphere_stokes(GPH, dP, lon, lat, **kw)
    assert_structure(Y.clm, Y.slm, L, L)
    e = rel_to_max(Y.clm, Y.slm, O.clm, O.slm)
    print('%s top %g km: %.2e' % (method, top_km, e))
    assert e <= TOL


def test_geopotential_instead_of_height_is_still_exact():
    """A caller passing geopotential (m^2/s^2, ~10x the height) lands far outside the expansion range."""
    from oracle import stokes_oracle
    from model_harmonics_b200 import testing as syn
    mh = _mh()
    L = 40
    lon, lat = syn.grid_axes(31, 60)
    GPH, dP, geoid = syn.atmosphere_cube(9, 31, 60, seed=32)
    kw = dict(LMAX=L, ELLIPSOID='WGS84', GEOID=geoid, LOVE=syn.love_numbers(L))
    Y = mh.gen_atmosphere_stokes(GPH * 9.80665, dP, lon, lat, **kw)
    O = stokes_oracle.gen_atmosphere_stokes(GPH * 9.80665, dP, lon, lat, **kw)
    assert rel_to_max(Y.clm, Y.slm, O.clm, O.slm) <= TOL


def test_guard_is_per_call_and_moment_and_direct_paths_agree(monkeypatch):
    from model_harmonics_b200 import testing as syn
    mh = _mh()
    L = 60
    lon, lat = syn.grid_axes(46, 90)
    GPH, dP, geoid = syn.atmosphere_cube(12, 46, 90, seed=33)
    kw = dict(LMAX=L, ELLIPSOID='WGS84', GEOID=geoid, LOVE=syn.love_numbers(L))
    a = mh.gen_atmosphere_stokes(GPH, dP, lon, lat, **kw)
    mh.gen_atmosphere_stokes(GPH * 5.0, dP, lon, lat, **kw)          # trips the guard
    b = mh.gen_atmosphere_stokes(GPH, dP, lon, lat, **kw)            # the flag does not leak into the next call
    assert np.array_equal(a.clm, b.clm) and np.array_equal(a.slm, b.slm)
    monkeypatch.setenv('MHB200_MOMENTS', '0')
    c = mh.gen_atmosphere_stokes(GPH, dP, lon, lat, **kw)
    assert rel_to_max(a.clm, a.slm, c.clm, c.slm) <= 1e-13
    assert not np.array_equal(a.clm, c.clm)                          # they really are different kernels


@pytest.mark.parametrize('L', [60, 200])
def test_pressure_radius_spread_inside_and_outside_the_guard(L):
    from oracle import stokes_oracle
    from model_harmonics_b200 import testing as syn
    mh = _mh()
    lon, lat = syn.grid_axes(61, 120)
    G, R = syn.surface_geometry(lon, lat, seed=34)
    P = syn.pressure_fields(61, 120, 1, seed=34)[0]
    LOVE = syn.love_numbers(L)
    rng = np.random.default_rng(35)
    # metres of extra random radius: 60 km is inside the guard at LMAX 60 (82 km) and outside at LMAX 200
    # (25 km); 200 km is outside at both
    for spread in (0.0, 60.0e3, 200.0e3):
        R2 = R + spread * rng.uniform(-1.0, 1.0, R.shape)
        Y = mh.gen_pressure_stokes(P, G, R2, lon, lat, LMAX=L, LOVE=LOVE)
        O = stokes_oracle.gen_pressure_stokes(P, G, R2, lon, lat, LMAX=L, LOVE=LOVE)
        e = rel_to_max(Y.clm, Y.slm, O.clm, O.slm)
        print('LMAX %d spread %g m: %.2e' % (L, spread, e))
        assert e <= TOL


def test_pressure_bad_radius_values_propagate_like_the_reference():
    """A NaN radius must give NaN coefficients (np.power propagates it), not a silently finite answer."""
    from model_harmonics_b200 import testing as syn
    mh = _mh()
    lon, lat = syn.grid_axes(31, 60)
    G, R = syn.surface_geometry(lon, lat, seed=36)
    P = syn.pressure_fields(31, 60, 1, seed=36)[0]
    R = R.copy()
    R[7, 11] = np.nan
    Y = mh.gen_pressure_stokes(P, G, R, lon, lat, LMAX=20, LOVE=syn.love_numbers(20))
    assert np.isnan(Y.clm).any()


def test_moment_path_runs_three_kernels_plus_two_guarded_stubs():
    import torch
    from model_harmonics_b200 import testing as syn
    mh = _mh()
    lon, lat = syn.grid_axes(46, 90)
    GPH, dP, geoid = syn.atmosphere_cube(5, 46, 90, seed=37)
    g, d = torch.from_numpy(GPH).cuda(), torch.from_numpy(dP).cuda()
    kw = dict(LMAX=30, ELLIPSOID='WGS84', GEOID=geoid, LOVE=syn.love_numbers(30), as_numpy=False)
    mh.atmosphere_stokes_batch(g[None], d[None], lon, lat, **kw)
    n0 = _launches()
    mh.atmosphere_stokes_batch(g[None], d[None], lon, lat, **kw)
    assert _launches() - n0 == 6          # ring constants, K_A, K_B, K_C + the direct ring / reduce kernels (exit at once)


def test_flipped_views_are_accepted_like_the_reference():
    """Negative-stride operands (P[::-1], GPH[:, ::-1]) -- ADVICE r1: torch.from_numpy rejects them."""
    from oracle import stokes_oracle
    from model_harmonics_b200 import testing as syn
    mh = _mh()
    L = 30
    LOVE = syn.love_numbers(L)
    lon, lat = syn.grid_axes(46, 90)
    G, R = syn.surface_geometry(lon, lat, seed=38)
    P = syn.pressure_fields(46, 90, 1, seed=38)[0]
    lat_r = lat[::-1]
    Y = mh.gen_pressure_stokes(P[::-1, :], G[::-1, :], R[::-1, :], lon, lat_r, LMAX=L, LOVE=LOVE)
    O = stokes_oracle.gen_pressure_stokes(P[::-1, :], G[::-1, :], R[::-1, :], lon, lat_r, LMAX=L, LOVE=LOVE)
    assert rel_to_max(Y.clm, Y.slm, O.clm, O.slm) <= TOL
    GPH, dP, geoid = syn.atmosphere_cube(4, 46, 90, seed=39)
    kw = dict(LMAX=L, ELLIPSOID='WGS84', LOVE=LOVE)
    Y = mh.gen_atmosphere_stokes(GPH[:, ::-1], dP[:, ::-1], lon, lat_r, GEOID=geoid[::-1], **kw)
    O = stokes_oracle.gen_atmosphere_stokes(GPH[:, ::-1], dP[:, ::-1], lon, lat_r, GEOID=geoid[::-1], **kw)
    assert rel_to_max(Y.clm, Y.slm, O.clm, O.slm) <= TOL


def test_plm_table_differing_in_unsampled_entries_is_honoured():
    """a caller table that differs from the standard one anywhere must be used as given."""
    from oracle import stokes_oracle
    from oracle.gravtk_standin import plm_holmes
    from model_harmonics_b200 import testing as syn
    mh = _mh()
    L = 40
    LOVE = syn.love_numbers(L)
    lon, lat = syn.grid_axes(61, 120)
    G, R = syn.surface_geometry(lon, lat, seed=40)
    P = syn.pressure_fields(61, 120, 1, seed=40)[0]
    PLM, _ = plm_holmes(L, np.cos(np.radians(90.0 - lat)))
    PLM2 = PLM.copy()
    PLM2[17, 5, :] = 0.0                  # one filtered (l, m) row: not among the entries round 1 sampled
    PLM2[33, 21, 30] *= 1.001
    Y = mh.gen_pressure_stokes(P, G, R, lon, lat, LMAX=L, PLM=PLM2, LOVE=LOVE)
    O = stokes_oracle.gen_pressure_stokes(P, G, R, lon, lat, LMAX=L, PLM=PLM2, LOVE=LOVE)
    assert rel_to_max(Y.clm, Y.slm, O.clm, O.slm) <= TOL
    assert Y.clm[17, 5] == 0.0 and Y.slm[17, 5] == 0.0
    # and the standard table still takes the on-the-fly path with identical results to PLM=None
    A = mh.gen_pressure_stokes(P, G, R, lon, lat, LMAX=L, PLM=PLM, LOVE=LOVE)
    B = mh.gen_pressure_stokes(P, G, R, lon, lat, LMAX=L, LOVE=LOVE)
    assert np.array_equal(A.clm, B.clm) and np.array_equal(A.slm, B.slm)


def test_two_threads_two_streams_share_one_plan():
    """SURVEY 8b re-entrancy: different fields in flight on two streams of one plan, from two threads."""
    import threading
    import torch
    from model_harmonics_b200 import testing as syn
    mh = _mh()
    L = 60
    lon, lat = syn.grid_axes(181, 360)
    LOVE = syn.love_numbers(L)
    G, R = syn.surface_geometry(lon, lat, seed=60)
    Gd, Rd = torch.from_numpy(G).cuda(), torch.from_numpy(R).cuda()
    P = [torch.from_numpy(syn.pressure_fields(181, 360, 6, seed=61 + 10 * i)).cuda() for i in range(2)]
    GPH, dP, geoid = syn.atmosphere_cube(20, 181, 360, seed=62)
    cubes = [(torch.from_numpy(GPH * (1.0 + 0.01 * i)).cuda(), torch.from_numpy(dP * (1.0 - 0.02 * i)).cuda())
             for i in range(2)]
    ge = torch.from_numpy(geoid).cuda()
    akw = dict(LMAX=L, ELLIPSOID='WGS84', GEOID=ge, LOVE=LOVE, as_numpy=False)
    pts = [[torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in syn.point_cloud(3000 + 500 * i, seed=70 + i)]
           for i in range(2)]

    def work(i):
        a = mh.pressure_stokes_batch(P[i], Gd, Rd, lon, lat, LMAX=L, LOVE=LOVE, as_numpy=False)
        b = mh.atmosphere_stokes_batch(cubes[i][0][None], cubes[i][1][None], lon, lat, **akw)
        pp = pts[i]
        c = mh.point_pressure_batch(pp[0][None], pp[1], pp[2], pp[3], pp[4], AREA=pp[5], LMAX=L, LOVE=LOVE,
                                    as_numpy=False)
        return [x.clone() for x in a + b + c]

    want = [work(0), work(1)]                      # single-threaded, default stream
    torch.cuda.synchronize()
    got = [None, None]
    errors = []

    def runner(i):
        try:
            torch.cuda.set_device(0)
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                for _ in range(25):                # many calls in flight on both streams at once
                    got[i] = work(i)
            stream.synchronize()
        except Exception as exc:                   # surfaced in the main thread
            errors.append(exc)

    threads = [threading.Thread(target=runner, args=(i,)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for i in range(2):
        for a, b in zip(got[i], want[i]):
            assert torch.equal(a, b)               # fixed-order reductions: bit-identical to the serial run


@pytest.mark.parametrize('nlon,lon0,centred,moment', [
    (8, 0.0, False, True), (12, -180.0, False, True), (64, 0.0, True, True), (132, -180.0, True, True),
    (260, 0.0, False, True), (1440, -180.0, False, True),
    # rotated by three steps: still symmetric, but the fold's base columns are no contiguous run -> direct kernels
    (516, 3 * 360.0 / 516, False, False),
    # arbitrary offset: no mirror symmetry at all -> direct kernels, unfolded
    (96, 7.5, False, False)])
def test_moment_path_on_other_regular_longitude_grids(nlon, lon0, centred, moment):
    """Grids starting at -180, cell-centred grids (no self-paired columns), grids rotated by whole steps, tiny and
    multi-chunk rings: fold pairing, TMA box starts (aligned below odd starts, negative starts) and the
    dropped duplicate columns, through both grid operators against the oracle."""
    import torch
    from oracle import stokes_oracle
    from model_harmonics_b200 import testing as syn
    mh = _mh()
    nlat, nlev, L = 9, 6, 20
    step = 360.0 / nlon
    lon = lon0 + (np.arange(nlon) + (0.5 if centred else 0.0)) * step
    lat = np.linspace(80.0, -80.0, nlat)
    GPH, dP, geoid = syn.atmosphere_cube(nlev, nlat, nlon, seed=nlon)
    LOVE = syn.love_numbers(L)
    kw = dict(LMAX=L, ELLIPSOID='WGS84', GEOID=geoid, LOVE=LOVE)
    n0 = _launches()
    c, s = mh.atmosphere_stokes_batch(torch.from_numpy(GPH).cuda()[None], torch.from_numpy(dP).cuda()[None], lon, lat,
                                      **kw)
    took_moment_path = (_launches() - n0 == 6)
    O = stokes_oracle.gen_atmosphere_stokes(GPH, dP, lon, lat, **kw)
    assert rel_to_max(c[0], s[0], O.clm, O.slm) <= TOL
    assert took_moment_path == moment
    G, R = syn.surface_geometry(np.mod(lon, 360.0), lat, seed=nlon)
    P = syn.pressure_fields(nlat, nlon, 1, seed=nlon)[0]
    Y = mh.gen_pressure_stokes(P, G, R, lon, lat, LMAX=L, MMAX=11, LOVE=LOVE)
    O = stokes_oracle.gen_pressure_stokes(P, G, R, lon, lat, LMAX=L, MMAX=11, LOVE=LOVE)
    assert rel_to_max(Y.clm, Y.slm, O.clm, O.slm) <= TOL


def test_sub_batching_of_long_batches():
    """More fields than one K_A launch takes (32): 35 distinct float32 cubes, 40 fields off a shared cube,
    37 pressure fields at LMAX 100; every field against its own single call."""
    import torch
    from model_harmonics_b200 import testing as syn
    mh = _mh()
    L = 30
    lon, lat = syn.grid_axes(19, 40)
    LOVE = syn.love_numbers(L)
    cubes = [syn.atmosphere_cube(9, 19, 40, seed=90 + t, dtype=np.float32) for t in range(35)]
    GPH = np.stack([c[0] for c in cubes])
    dP = np.stack([c[1] for c in cubes])
    kw = dict(LMAX=L, ELLIPSOID='WGS84', GEOID=cubes[0][2], LOVE=LOVE)
    n0 = _launches()
    clm, slm = mh.atmosphere_stokes_batch(torch.from_numpy(GPH).cuda(), torch.from_numpy(dP).cuda(), lon, lat, **kw)
    assert _launches() - n0 == 1 + 2 * 3 + 2
    for t in (0, 31, 32, 34):
        c1, s1 = mh.atmosphere_stokes_batch(torch.from_numpy(GPH[t:t + 1]).cuda(), torch.from_numpy(dP[t:t + 1]).cuda(),
                                            lon, lat, **kw)
        assert rel_to_max(clm[t], slm[t], c1[0], s1[0]) <= 1e-14
    scales = np.array([syn.atmosphere_field_scalars(t) for t in range(40)])
    g64, d64 = GPH[0].astype(np.float64), dP[0].astype(np.float64)
    clm, slm = mh.atmosphere_stokes_batch(g64, d64, lon, lat, field_scale=scales, **kw)
    for t in (0, 32, 39):
        Y = mh.gen_atmosphere_stokes(g64 * scales[t, 0], d64 * scales[t, 1], lon, lat, **kw)
        assert rel_to_max(clm[t], slm[t], Y.clm, Y.slm) <= 1e-14
    G, R = syn.surface_geometry(lon, lat, seed=91)
    P = syn.pressure_fields(19, 40, 37, seed=92)
    LV = syn.love_numbers(100)
    clm, slm = mh.pressure_stokes_batch(P, G, R, lon, lat, LMAX=100, LOVE=LV)
    for t in (0, 31, 36):
        Y = mh.gen_pressure_stokes(P[t], G, R, lon, lat, LMAX=100, LOVE=LV)
        assert rel_to_max(clm[t], slm[t], Y.clm, Y.slm) <= 1e-14


def test_randomised_differential_moment_vs_direct_kernels(monkeypatch):
    """Forty random problem shapes (rings, longitudes, levels, LMAX/MMAX, method, storage type, batch size,
    shared cube or distinct cubes) through the contracted-moment kernels and through the direct kernels
    (MHB200_MOMENTS=0): two independent implementations of the same sums must agree to rounding.  Covers slot
    bookkeeping with odd ring / batch / chunk counts, K_B ring splits, tail level groups and partial chunks."""
    import torch
    from model_harmonics_b200 import testing as syn
    mh = _mh()
    rng = np.random.default_rng(2025)
    worst = 0.0
    for case in range(40):
        nlon = 4 * int(rng.integers(2, 140))
        nlat = int(rng.integers(2, 40))
        L = int(rng.integers(0, 64))
        M = None if rng.uniform() < 0.6 else int(rng.integers(0, L + 1))
        LOVE = syn.love_numbers(L)
        lon = (-180.0 if rng.uniform() < 0.3 else 0.0) + np.arange(nlon) * (360.0 / nlon)
        lat = np.linspace(90.0, -90.0, nlat) if rng.uniform() < 0.5 else np.linspace(85.0, -85.0, nlat)
        T = int(rng.integers(1, 12))
        if rng.uniform() < 0.55:
            nlev = int(rng.integers(1, 23))
            method = 'BC05' if rng.uniform() < 0.6 else 'SW02'
            dtype = np.float32 if (rng.uniform() < 0.3 and nlon % 8 == 0) else np.float64
            GPH, dP, geoid = syn.atmosphere_cube(nlev, nlat, nlon, seed=1000 + case, dtype=dtype)
            kw = dict(LMAX=L, MMAX=M, ELLIPSOID='WGS84', GEOID=geoid, LOVE=LOVE, METHOD=method)
            if rng.uniform() < 0.5:
                scales = 1.0 + 1e-3 * rng.standard_normal((T, 2))
                call = lambda: mh.atmosphere_stokes_batch(GPH, dP, lon, lat, field_scale=scales, **kw)
            else:
                g4 = np.stack([GPH * dtype(1.0 + 0.01 * t) for t in range(T)])
                d4 = np.stack([dP * dtype(1.0 - 0.02 * t) for t in range(T)])
                call = lambda: mh.atmosphere_stokes_batch(torch.from_numpy(g4).cuda(), torch.from_numpy(d4).cuda(),
                                                          lon, lat, **kw)
        else:
            G, R = syn.surface_geometry(np.mod(lon, 360.0), lat, seed=1000 + case)
            P = syn.pressure_fields(nlat, nlon, T, seed=2000 + case, mean=float(rng.choice([0.0, 1.0e5])))
            call = lambda: mh.pressure_stokes_batch(P, G, R, lon, lat, LMAX=L, MMAX=M, LOVE=LOVE)
        monkeypatch.setenv('MHB200_MOMENTS', '2')
        a = call()
        monkeypatch.setenv('MHB200_MOMENTS', '0')
        b = call()
        for t in range(T):
            scale = max(np.abs(b[0][t]).max(), np.abs(b[1][t]).max())
            e = max(np.abs(a[0][t] - b[0][t]).max(), np.abs(a[1][t] - b[1][t]).max()) / scale
            worst = max(worst, e)
            assert e <= 1e-13, (case, t, nlat, nlon, L, M, e)
    print('randomised differential test: worst relative-to-max difference %.2e' % worst)
