"""
GPU parity at the awkward sizes: tile and chunk boundaries, multi-group level loops, tiny grids,
single points, high degree; plus oracle-free properties at the BASELINE sizes.
"""
import numpy as np
import pytest

from parity import TOL, rel_to_max, assert_structure

pytestmark = pytest.mark.gpu


def _mh():
    import model_harmonics_b200 as mh
    return mh


def _syn():
    from model_harmonics_b200 import testing as syn
    return syn


@pytest.mark.parametrize('nlat,nlon,L,M', [
    (2, 4, 0, None), (3, 7, 1, None), (5, 9, 2, 1), (19, 37, 63, None), (19, 36, 64, None),
    (21, 40, 65, 63), (17, 33, 70, 64), (13, 130, 20, None), (11, 257, 12, 5), (9, 129, 130, 3),
])
def test_pressure_boundary_sizes(nlat, nlon, L, M):
    from oracle import stokes_oracle
    mh, syn = _mh(), _syn()
    rng = np.random.default_rng(nlat * 1000 + nlon)
    lat = np.linspace(88.0, -88.0, nlat)
    lon = np.arange(nlon) * (360.0 / nlon)
    P = 500.0 * rng.standard_normal((nlat, nlon))
    G = 9.8 + 0.01 * rng.standard_normal((nlat, nlon))
    R = 6.371e6 + 2.0e3 * rng.standard_normal((nlat, nlon))
    kw = dict(LMAX=L, MMAX=M, LOVE=syn.love_numbers(L))
    A = mh.gen_pressure_stokes(P, G, R, lon, lat, **kw)
    B = stokes_oracle.gen_pressure_stokes(P, G, R, lon, lat, **kw)
    assert_structure(A.clm, A.slm, int(B.lmax), int(B.mmax))
    assert rel_to_max(A.clm, A.slm, B.clm, B.slm) <= TOL


@pytest.mark.parametrize('nlev,nlat,nlon,L,M,method', [
    (1, 7, 12, 5, None, 'BC05'), (2, 9, 16, 8, 3, 'SW02'), (41, 9, 20, 10, None, 'BC05'),
    (80, 7, 16, 6, None, 'BC05'), (81, 7, 16, 6, None, 'SW02'), (5, 11, 70, 66, None, 'BC05'),
    (3, 9, 131, 70, 20, 'SW02'),
])
def test_atmosphere_boundary_sizes(nlev, nlat, nlon, L, M, method):
    from oracle import stokes_oracle
    mh, syn = _mh(), _syn()
    lat = np.linspace(90.0, -90.0, nlat)
    lon = np.arange(nlon) * (360.0 / nlon)
    GPH, dP, geoid = syn.atmosphere_cube(nlev, nlat, nlon, seed=nlev + nlon)
    kw = dict(LMAX=L, MMAX=M, ELLIPSOID='WGS84', GEOID=geoid, LOVE=syn.love_numbers(L), METHOD=method)
    A = mh.gen_atmosphere_stokes(GPH, dP, lon, lat, **kw)
    B = stokes_oracle.gen_atmosphere_stokes(GPH, dP, lon, lat, **kw)
    assert_structure(A.clm, A.slm, int(B.lmax), int(B.mmax))
    assert rel_to_max(A.clm, A.slm, B.clm, B.slm) <= TOL
    # scalar geoid broadcasts (reference: N + GEOID)
    kw['GEOID'] = 12.5
    A = mh.gen_atmosphere_stokes(GPH, dP, lon, lat, **kw)
    B = stokes_oracle.gen_atmosphere_stokes(GPH, dP, lon, lat, **kw)
    assert rel_to_max(A.clm, A.slm, B.clm, B.slm) <= TOL


@pytest.mark.parametrize('n,L,M', [(1, 4, None), (31, 0, None), (33, 1, None), (257, 70, None), (100, 100, 64),
                                    (64, 12, 0), (5000, 130, 101), (20000, 61, 3)])
def test_points_boundary_sizes(n, L, M):
    from oracle import stokes_oracle
    mh, syn = _mh(), _syn()
    P, G, R, lon, lat, AREA = syn.point_cloud(n, seed=n + L, pole_points=(n >= 8))
    kw = dict(AREA=AREA, LMAX=L, MMAX=M, LOVE=syn.love_numbers(L))
    A = mh.gen_point_pressure(P, G, R, lon, lat, **kw)
    B = stokes_oracle.gen_point_pressure(P, G, R, lon, lat, **kw)
    assert A.clm.shape == B.clm.shape
    assert rel_to_max(A.clm, A.slm, B.clm, B.slm) <= TOL


@pytest.mark.parametrize('n,L,M', [(4097, 60, None), (1000, 61, 40), (129, 33, 33), (70000, 60, 60), (63, 7, 5)])
def test_points_fused_and_staged_kernels_agree(monkeypatch, n, L, M):
    """The fused kernel (one launch, one-multiplier recursion, in-kernel weights and trig factors; default for
    LMAX <= 61) against the round-1 prepare + contract kernels (MHB200_POINTS_FUSED=0), both against the oracle."""
    from oracle import stokes_oracle
    mh, syn = _mh(), _syn()
    P, G, R, lon, lat, AREA = syn.point_cloud(n, seed=3 * n + L, pole_points=True)
    kw = dict(AREA=AREA, LMAX=L, MMAX=M, LOVE=syn.love_numbers(L))
    ref = stokes_oracle.gen_point_pressure(P, G, R, lon, lat, **kw) if n <= 5000 else None
    out = {}
    for flag in ('1', '0'):
        monkeypatch.setenv('MHB200_POINTS_FUSED', flag)
        out[flag] = mh.gen_point_pressure(P, G, R, lon, lat, **kw)
        if ref is not None:
            assert rel_to_max(out[flag].clm, out[flag].slm, ref.clm, ref.slm) <= TOL
    monkeypatch.delenv('MHB200_POINTS_FUSED')
    assert rel_to_max(out['0'].clm, out['0'].slm, out['1'].clm, out['1'].slm) <= 1e-13
    assert np.abs(out['0'].clm - out['1'].clm).max() > 0.0        # the two paths really are different code
    # bitwise repeatable
    again = mh.gen_point_pressure(P, G, R, lon, lat, **kw)
    assert np.array_equal(again.clm, out['1'].clm) and np.array_equal(again.slm, out['1'].slm)


def test_points_randomised_differential_fused_vs_staged(monkeypatch):
    """Forty random point problems (count from 1 to a few thousand, LMAX 0..61, MMAX, batch size, scalar or
    per-point G / R, poles included) through the fused kernel and through the prepare + contract kernels: two
    independent implementations of the same sums must agree to rounding, field by field."""
    mh, syn = _mh(), _syn()
    rng = np.random.default_rng(4242)
    worst = 0.0
    for case in range(40):
        n = int(rng.choice([1, 2, 63, 64, 65, 127, 129, int(rng.integers(3, 6000))]))
        L = int(rng.integers(0, 62))
        M = None if rng.uniform() < 0.5 else int(rng.integers(0, L + 1))
        T = int(rng.integers(1, 5))
        P, G, R, lon, lat, AREA = syn.point_cloud(n, seed=9000 + case, pole_points=bool(rng.uniform() < 0.5))
        Pb = np.stack([P * (1.0 + 0.1 * t) + 3.0 * t for t in range(T)])
        if rng.uniform() < 0.3:
            G = float(np.mean(G))                  # scalar operands broadcast (reference: G.flatten() of a scalar)
        kw = dict(AREA=AREA, LMAX=L, MMAX=M, LOVE=syn.love_numbers(L))
        out = {}
        for flag in ('1', '0'):
            monkeypatch.setenv('MHB200_POINTS_FUSED', flag)
            out[flag] = mh.point_pressure_batch(Pb, G, R, lon, lat, **kw)
        for t in range(T):
            a, b = out['1'], out['0']
            scale = max(np.abs(b[0][t]).max(), np.abs(b[1][t]).max())
            e = max(np.abs(a[0][t] - b[0][t]).max(), np.abs(a[1][t] - b[1][t]).max()) / scale
            worst = max(worst, e)
            assert e <= 1e-13, (case, t, n, L, M, e)
    monkeypatch.delenv('MHB200_POINTS_FUSED')
    print('randomised point differential test: worst relative-to-max difference %.2e' % worst)


@pytest.mark.parametrize('n', [64, 65, 6400, 100000])
def test_points_fused_kernel_is_bitwise_repeatable(n):
    """The fused kernel hands its tile buffers over with named barriers between producer and contraction warps; a
    missed hand-over would show as run-to-run differences.  Sixty calls, device-resident, must agree bit for bit."""
    import torch
    mh, syn = _mh(), _syn()
    P, G, R, lon, lat, AREA = syn.point_cloud(n, seed=n)
    dv = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (P, G, R, lon, lat, AREA)]
    kw = dict(AREA=dv[5], LMAX=60, LOVE=syn.love_numbers(60), as_numpy=False)
    c0, s0 = mh.point_pressure_batch(dv[0][None], dv[1], dv[2], dv[3], dv[4], **kw)
    c0, s0 = c0.clone(), s0.clone()
    for _ in range(60):
        c, s = mh.point_pressure_batch(dv[0][None], dv[1], dv[2], dv[3], dv[4], **kw)
        assert torch.equal(c, c0) and torch.equal(s, s0)


def test_short_sincos_of_the_fused_point_kernel_is_within_two_ulp():
    """The fused kernel's own sincos (Cody-Waite with a three-part pi/2, Taylor polynomials) against NumPy's on
    two million arguments up to its range limit: quadrant boundaries, tiny and huge-ish arguments included."""
    import torch
    from model_harmonics_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(7)
    a = np.concatenate([rng.uniform(-7.0, 7.0, 500000), rng.uniform(-400.0, 400.0, 500000),
                        rng.uniform(-4.9e5, 4.9e5, 900000),
                        (np.arange(-50000, 50000) * (np.pi / 4)) * (1.0 + 1e-16 * rng.integers(-4, 5, 100000)),
                        np.array([0.0, -0.0, 1e-300, -1e-300, 1e-8, 4.99e5, -4.99e5])])
    at = torch.from_numpy(a).cuda()
    st, ct = torch.empty_like(at), torch.empty_like(at)
    rc = L.mhb200_sincos_probe(0, at.data_ptr(), a.size, st.data_ptr(), ct.data_ptr(), None)
    _lib.check(rc, 'mhb200_sincos_probe')
    torch.cuda.synchronize()
    s, c = st.cpu().numpy(), ct.cpu().numpy()
    es = np.abs(s - np.sin(a)) / np.spacing(np.maximum(np.abs(np.sin(a)), 1e-300))
    ec = np.abs(c - np.cos(a)) / np.spacing(np.maximum(np.abs(np.cos(a)), 1e-300))
    # near a zero of sin / cos the error is bounded absolutely (the reduction carries ~1e-17), not in ulps of the result
    ok_s = (es <= 2.0) | (np.abs(s - np.sin(a)) <= 2e-16)
    ok_c = (ec <= 2.0) | (np.abs(c - np.cos(a)) <= 2e-16)
    assert ok_s.all() and ok_c.all(), (float(es[~ok_s].max()) if (~ok_s).any() else 0.0,
                                       float(ec[~ok_c].max()) if (~ok_c).any() else 0.0)
    print('short sincos: max error %.2f / %.2f ulp where the result is not near zero' % (
        es[np.abs(np.sin(a)) > 0.1].max(), ec[np.abs(np.cos(a)) > 0.1].max()))


def test_points_many_turns_of_longitude_stay_exact():
    """Longitudes a thousand turns away: m phi ~ 4e5 is still inside the fused kernel's short sincos (Cody-Waite with
    k < 2^19); the result must follow the reference's exp(1j m phi) of the rounded product."""
    from oracle import stokes_oracle
    mh, syn = _mh(), _syn()
    P, G, R, lon, lat, AREA = syn.point_cloud(900, seed=11)
    lon = lon - 360.0 * 1000.0
    kw = dict(AREA=AREA, LMAX=60, MMAX=None, LOVE=syn.love_numbers(60))
    A = mh.gen_point_pressure(P, G, R, lon, lat, **kw)
    B = stokes_oracle.gen_point_pressure(P, G, R, lon, lat, **kw)
    assert rel_to_max(A.clm, A.slm, B.clm, B.slm) <= TOL


def test_points_huge_angles_take_the_safe_kernel():
    """Longitudes / colatitudes far outside a circle (|m phi| beyond the fused kernel's short sincos) are redone by
    the general-purpose variant of the same kernel: still the reference's exp(1j m phi) of the rounded product."""
    from oracle import stokes_oracle
    mh, syn = _mh(), _syn()
    P, G, R, lon, lat, AREA = syn.point_cloud(700, seed=5)
    lon = lon + 360.0 * 2.0e5                     # phi ~ 1.3e6 rad
    kw = dict(AREA=AREA, LMAX=40, MMAX=None, LOVE=syn.love_numbers(40))
    A = mh.gen_point_pressure(P, G, R, lon, lat, **kw)
    B = stokes_oracle.gen_point_pressure(P, G, R, lon, lat, **kw)
    assert rel_to_max(A.clm, A.slm, B.clm, B.slm) <= TOL
    # and a normal call afterwards is unaffected
    P, G, R, lon, lat, AREA = syn.point_cloud(700, seed=6)
    A = mh.gen_point_pressure(P, G, R, lon, lat, **kw)
    B = stokes_oracle.gen_point_pressure(P, G, R, lon, lat, **kw)
    assert rel_to_max(A.clm, A.slm, B.clm, B.slm) <= TOL


@pytest.mark.parametrize('method', ['BC05', 'SW02'])
def test_moment_path_and_direct_path_agree(monkeypatch, method):
    """The binomial-moment produce path (default at fold level 2, LMAX <= 63) against the direct powers, both
    against the oracle; 45 levels reaching 80 km also exercise two level groups and large |r/r0 - 1|."""
    from oracle import stokes_oracle
    mh, syn = _mh(), _syn()
    nlev, nlat, nlon, L = 45, 19, 48, 60
    lat = np.linspace(90.0, -90.0, nlat)
    lon = np.arange(nlon) * (360.0 / nlon)
    GPH, dP, geoid = syn.atmosphere_cube(nlev, nlat, nlon, seed=77)
    GPH = GPH * (80000.0 / GPH.max())            # top level at 80 km
    kw = dict(LMAX=L, ELLIPSOID='WGS84', GEOID=geoid, LOVE=syn.love_numbers(L), METHOD=method)
    ref = stokes_oracle.gen_atmosphere_stokes(GPH, dP, lon, lat, **kw)
    out = {}
    for flag in ('0', '1'):
        monkeypatch.setenv('MHB200_MOMENTS', flag)
        out[flag] = mh.gen_atmosphere_stokes(GPH, dP, lon, lat, **kw)
        assert rel_to_max(out[flag].clm, out[flag].slm, ref.clm, ref.slm) <= TOL
    monkeypatch.delenv('MHB200_MOMENTS')
    assert rel_to_max(out['0'].clm, out['0'].slm, out['1'].clm, out['1'].slm) <= 1e-13
    assert np.abs(out['0'].clm - out['1'].clm).max() > 0.0        # the two paths really are different code


def test_pressure_lmax360_against_oracle():
    """High degree (21 tiles, FULL + DIAG paths, deep Legendre recursion incl. polar underflow)."""
    from oracle import stokes_oracle
    mh, syn = _mh(), _syn()
    lon, lat = syn.grid_axes(91, 180)
    G, R = syn.surface_geometry(lon, lat, seed=4)
    P = syn.pressure_fields(91, 180, 1, seed=4)[0]
    kw = dict(LMAX=360, LOVE=syn.love_numbers(360))
    A = mh.gen_pressure_stokes(P, G, R, lon, lat, **kw)
    B = stokes_oracle.gen_pressure_stokes(P, G, R, lon, lat, **kw)
    assert_structure(A.clm, A.slm, 360, 360)
    assert rel_to_max(A.clm, A.slm, B.clm, B.slm) <= TOL


def test_full_size_properties_pressure_c4():
    """0.25 degree, LMAX=360, no oracle: linearity, longitude-shift phase rotation, zonal field."""
    import torch
    mh, syn = _mh(), _syn()
    nlat, nlon, L = 721, 1440, 360
    lon, lat = syn.grid_axes(nlat, nlon)
    LOVE = syn.love_numbers(L)
    rng = np.random.default_rng(77)
    P1 = torch.from_numpy(500.0 * rng.standard_normal((nlat, nlon))).cuda()
    P2 = torch.from_numpy(500.0 * rng.standard_normal((nlat, nlon))).cuda()
    k = 37                                           # shift by 37 cells in longitude
    zonal = torch.from_numpy(np.repeat(300.0 * rng.standard_normal((nlat, 1)), nlon, axis=1)).cuda()
    batch = torch.stack([P1, P2, 2.0 * P1 - 0.5 * P2, torch.roll(P1, k, dims=1), zonal])
    # constant G and a radius that depends on latitude only keep the shift property exact
    Rlat = torch.from_numpy(6.371e6 + 1.0e4 * np.cos(np.radians(lat))[:, None] * np.ones((1, nlon))).cuda()
    clm, slm = mh.pressure_stokes_batch(batch, 9.81, Rlat, lon, lat, LMAX=L, LOVE=LOVE)
    scale = np.abs(clm[0]).max()
    # linearity
    assert np.abs(2.0 * clm[0] - 0.5 * clm[1] - clm[2]).max() <= 1e-12 * scale
    assert np.abs(2.0 * slm[0] - 0.5 * slm[1] - slm[2]).max() <= 1e-12 * scale
    # field shifted east by k cells: (C + iS)_m -> (C + iS)_m * exp(i m k dphi)
    ang = np.arange(L + 1) * (k * 2.0 * np.pi / nlon)
    z = (clm[0] + 1j * slm[0]) * np.exp(1j * ang)[None, :]
    assert np.abs(z.real - clm[3]).max() <= 1e-11 * scale
    assert np.abs(z.imag - slm[3]).max() <= 1e-11 * scale
    # zonal field: only order 0
    zs = np.abs(clm[4][:, 0]).max()
    assert np.abs(clm[4][:, 1:]).max() <= 1e-12 * zs and np.abs(slm[4]).max() <= 1e-12 * zs
    for t in range(5):
        assert_structure(clm[t], slm[t], L, L)


def test_full_size_properties_atmosphere_c3():
    """ERA5-shaped cube: dP linearity and sign, batch scale factors, against a latitude sub-sampled oracle run."""
    import torch
    from oracle import stokes_oracle
    mh, syn = _mh(), _syn()
    nlev, nlat, nlon, L = 37, 721, 1440, 60
    lon, lat = syn.grid_axes(nlat, nlon)
    GPH, dP, geoid = syn.atmosphere_cube(nlev, nlat, nlon, seed=1234)
    LOVE = syn.love_numbers(L)
    g, d = torch.from_numpy(GPH).cuda(), torch.from_numpy(dP).cuda()
    kw = dict(LMAX=L, ELLIPSOID='WGS84', GEOID=torch.from_numpy(geoid).cuda(), LOVE=LOVE)
    sc = np.array([[1.0, 1.0], [1.0, -2.0], [1.0, 0.25]])
    clm, slm = mh.atmosphere_stokes_batch(g, d, lon, lat, field_scale=sc, **kw)
    scale = np.abs(clm[0]).max()
    # the transform is linear in dP for fixed GPH
    assert np.abs(clm[1] + 2.0 * clm[0]).max() <= 1e-13 * scale
    assert np.abs(slm[2] - 0.25 * slm[0]).max() <= 1e-13 * scale
    # rings are additive: rings 0::8 with explicit weights, CUDA vs oracle (a few seconds of CPU)
    sub = slice(0, nlat, 8)
    th = np.radians(90.0 - lat[sub])
    w = np.sin(th) * np.radians(0.25) * np.radians(0.25)
    kw2 = dict(LMAX=L, ELLIPSOID='WGS84', GEOID=geoid[sub], WEIGHT=w, LOVE=LOVE)
    A = mh.gen_atmosphere_stokes(GPH[:, sub], dP[:, sub], lon, lat[sub], **kw2)
    B = stokes_oracle.gen_atmosphere_stokes(GPH[:, sub], dP[:, sub], lon, lat[sub], **kw2)
    assert rel_to_max(A.clm, A.slm, B.clm, B.slm) <= TOL
