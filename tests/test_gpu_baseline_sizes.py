"""
Parity at the BASELINE.json sizes (GPU): configs C2 (1e5 points, LMAX 60), C3 (37 x 721 x 1440 atmosphere,
LMAX 60, BC05 and SW02), C4 (721 x 1440 pressure, LMAX 360) and the C5 sample fields t in {0, 7, 123, 479},
each against a committed golden produced by the UNMODIFIED reference files
(``tests/golden/make_golden.py --big``).  Inputs are rebuilt from the same seeds and checked by SHA-256
of the golden's record.  Gate: max|delta| / max|ref| <= 1e-12.
"""
import functools
import os

import numpy as np
import pytest

import golden_cases
from parity import TOL, rel_to_max, assert_structure

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _gold(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


@functools.lru_cache(maxsize=1)
def _c3_inputs():
    return golden_cases.build_big('atmosphere_c3_bc05')


def _literal(Y, gold):
    """Informational: worst literal per-coefficient relative error over |ref| > 1e-3 max."""
    ref = np.concatenate([gold['clm'].ravel(), gold['slm'].ravel()])
    got = np.concatenate([Y.clm.ravel(), Y.slm.ravel()])
    big = np.abs(ref) > 1e-3 * np.abs(ref).max()
    return float((np.abs(got - ref)[big] / np.abs(ref)[big]).max())


def _per_degree(Y, gold):
    """Informational: max over degrees of max_m |delta_lm| / amplitude_l(ref)."""
    amp = np.sqrt((gold['clm']**2 + gold['slm']**2).sum(axis=1))
    err = np.maximum(np.abs(Y.clm - gold['clm']), np.abs(Y.slm - gold['slm'])).max(axis=1)
    ok = amp > 0
    return float((err[ok] / amp[ok]).max())


def test_c2_points_1e5_vs_reference_golden():
    import model_harmonics_b200 as mh
    func, args, kwargs = golden_cases.build_big('points_c2')
    gold = _gold('points_c2')
    assert golden_cases.digest(args, kwargs) == str(gold['input_sha256'])
    Y = mh.gen_point_pressure(*args, **kwargs)
    assert_structure(Y.clm, Y.slm, 60, 60)
    e = rel_to_max(Y.clm, Y.slm, gold['clm'], gold['slm'])
    print('C2 1e5 points: rel-to-max %.2e, literal per-coefficient (|ref| > 1e-3 max) %.2e, per-degree %.2e'
          % (e, _literal(Y, gold), _per_degree(Y, gold)))
    assert e <= TOL


@pytest.mark.parametrize('method', ['BC05', 'SW02'])
@pytest.mark.parametrize('moments', ['0', '1', '2'])
def test_c3_full_field_vs_reference_golden(method, moments, monkeypatch):
    import torch
    import model_harmonics_b200 as mh
    func, args, kwargs = _c3_inputs()
    kwargs = dict(kwargs, METHOD=method)
    gold = _gold('atmosphere_c3_' + method.lower())
    if method == 'BC05':
        assert golden_cases.digest(args, kwargs) == str(gold['input_sha256'])
    monkeypatch.setenv('MHB200_MOMENTS', moments)
    # device-resident cubes: all 721 rings in one launch
    g = torch.from_numpy(args[0]).cuda()
    d = torch.from_numpy(args[1]).cuda()
    Y = mh.gen_atmosphere_stokes(g, d, args[2], args[3], **kwargs)
    assert_structure(Y.clm, Y.slm, 60, 60)
    e = rel_to_max(Y.clm, Y.slm, gold['clm'], gold['slm'])
    print('C3 %s MHB200_MOMENTS=%s: rel-to-max %.2e, per-degree (error / degree amplitude) %.2e'
          % (method, moments, e, _per_degree(Y, gold)))
    assert e <= TOL


def test_c3_host_band_path_vs_reference_golden():
    import model_harmonics_b200 as mh
    func, args, kwargs = _c3_inputs()
    gold = _gold('atmosphere_c3_bc05')
    Y = mh.gen_atmosphere_stokes(*args, **kwargs)          # NumPy cubes: latitude-band H2D pipeline
    e = rel_to_max(Y.clm, Y.slm, gold['clm'], gold['slm'])
    print('C3 BC05 host-band path: rel-to-max %.2e' % e)
    assert e <= TOL


def test_c5_sample_fields_vs_reference_golden():
    """Fields t of the 480-field job are derived on load from the base cube."""
    import torch
    import model_harmonics_b200 as mh
    from model_harmonics_b200 import testing as syn
    func, args, kwargs = _c3_inputs()
    ts = (0, 7, 123, 479)
    scales = np.array([syn.atmosphere_field_scalars(t) for t in ts])
    g = torch.from_numpy(args[0]).cuda()
    d = torch.from_numpy(args[1]).cuda()
    kw = {k: v for k, v in kwargs.items() if k != 'METHOD'}
    clm, slm = mh.atmosphere_stokes_batch(g, d, args[2], args[3], field_scale=scales, METHOD='BC05', **kw)
    for i, t in enumerate(ts):
        gold = _gold('atmosphere_c5_t%d' % t)
        e = rel_to_max(clm[i], slm[i], gold['clm'], gold['slm'])
        print('C5 t=%d: rel-to-max %.2e' % (t, e))
        assert e <= TOL


def test_c4_pressure_lmax360_vs_reference_golden():
    import model_harmonics_b200 as mh
    func, args, kwargs = golden_cases.build_big('pressure_c4')
    gold = _gold('pressure_c4')
    assert golden_cases.digest(args, kwargs) == str(gold['input_sha256'])
    Y = mh.gen_pressure_stokes(*args, **kwargs)
    assert_structure(Y.clm, Y.slm, 360, 360)
    e = rel_to_max(Y.clm, Y.slm, gold['clm'], gold['slm'])
    print('C4 721x1440 LMAX 360: rel-to-max %.2e, literal %.2e, per-degree %.2e' % (e, _literal(Y, gold), _per_degree(Y, gold)))
    assert e <= TOL


def test_distinct_cube_batch_and_point_batch_vs_oracle():
    """4-D (T, level, lat, lon) cubes and a T = 3 point batch against per-field oracle calls."""
    from oracle import stokes_oracle
    import model_harmonics_b200 as mh
    from model_harmonics_b200 import testing as syn
    L = 40
    LOVE = syn.love_numbers(L)
    lon, lat = syn.grid_axes(61, 120)
    cubes = [syn.atmosphere_cube(7, 61, 120, seed=40 + t) for t in range(3)]
    GPH = np.stack([c[0] for c in cubes])
    dP = np.stack([c[1] for c in cubes])
    geoid = cubes[0][2]
    kw = dict(LMAX=L, ELLIPSOID='WGS84', GEOID=geoid, LOVE=LOVE)
    for method in ('BC05', 'SW02'):
        clm, slm = mh.atmosphere_stokes_batch(GPH, dP, lon, lat, METHOD=method, **kw)
        for t in range(3):
            O = stokes_oracle.gen_atmosphere_stokes(GPH[t], dP[t], lon, lat, METHOD=method, **kw)
            e = rel_to_max(clm[t], slm[t], O.clm, O.slm)
            assert e <= TOL, (method, t, e)
    P, G, R, plon, plat, AREA = syn.point_cloud(4000, seed=9)
    Pb = np.stack([P, 2.0 * P[::-1].copy(), P + 100.0])
    clm, slm = mh.point_pressure_batch(Pb, G, R, plon, plat, AREA=AREA, LMAX=L, LOVE=LOVE)
    for t in range(3):
        O = stokes_oracle.gen_point_pressure(Pb[t], G, R, plon, plat, AREA=AREA, LMAX=L, LOVE=LOVE)
        e = rel_to_max(clm[t], slm[t], O.clm, O.slm)
        assert e <= TOL, ('points', t, e)
