"""Quick device-resident timing of the three operators at the BASELINE configs (scratch tool)."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, '.')
import model_harmonics_b200 as mh
from model_harmonics_b200 import testing as syn

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))

which = sys.argv[1:] or ['c1', 'c2', 'c3', 'c4']
LOVE = syn.love_numbers(60)
if 'c3' in which:
    lon, lat = syn.grid_axes(721, 1440)
    GPH, dP, geoid = syn.atmosphere_cube(37, 721, 1440)
    g, d, ge = torch.from_numpy(GPH).cuda(), torch.from_numpy(dP).cuda(), torch.from_numpy(geoid).cuda()
    kw = dict(LMAX=60, ELLIPSOID='WGS84', GEOID=ge, LOVE=LOVE)
    f = lambda: mh.atmosphere_stokes_batch(g[None], d[None], lon, lat, as_numpy=False, **kw)
    med, best = timeit(f)
    print(json.dumps({'cfg': 'c3 atm 37x721x1440 L60', 'ms_med': med, 'ms_best': best, 'tflops_alg': 16.9e9 / best / 1e9}))
    sc = np.array([syn.atmosphere_field_scalars(t) for t in range(8)])
    f = lambda: mh.atmosphere_stokes_batch(g, d, lon, lat, field_scale=sc, as_numpy=False, **kw)
    med, best = timeit(f, n=5)
    print(json.dumps({'cfg': 'c5-like 8 fields', 'ms_med': med, 'ms_per_field': best / 8, 'tflops_alg': 8 * 16.9e9 / best / 1e9}))
    g32, d32 = g.float(), d.float()
    f = lambda: mh.atmosphere_stokes_batch(g32[None], d32[None], lon, lat, as_numpy=False, **kw)
    med, best = timeit(f)
    print(json.dumps({'cfg': 'c3 f32 cubes', 'ms_med': med, 'ms_best': best}))
    f = lambda: mh.atmosphere_stokes_batch(g[None], d[None], lon, lat, as_numpy=False, METHOD='SW02', **kw)
    med, best = timeit(f)
    print(json.dumps({'cfg': 'c3 SW02', 'ms_med': med, 'ms_best': best}))
    del g, d, g32, d32
if 'c1' in which:
    lon, lat = syn.grid_axes(181, 360)
    G, R = syn.surface_geometry(lon, lat)
    P = syn.pressure_fields(181, 360, 12)
    Pd, Gd, Rd = [torch.from_numpy(a).cuda() for a in (P, G, R)]
    f = lambda: mh.pressure_stokes_batch(Pd, Gd, Rd, lon, lat, LMAX=60, LOVE=LOVE, as_numpy=False)
    med, best = timeit(f, n=20)
    print(json.dumps({'cfg': 'c1 12x181x360 L60', 'ms_med': med, 'ms_best': best, 'tflops_alg': 12 * 0.5e9 / best / 1e9}))
    f = lambda: mh.pressure_stokes_batch(Pd[:1], Gd, Rd, lon, lat, LMAX=60, LOVE=LOVE, as_numpy=False)
    med, best = timeit(f, n=20)
    print(json.dumps({'cfg': 'c1 single field', 'ms_med': med, 'ms_best': best}))
if 'c2' in which:
    P, G, R, lon, lat, AREA = syn.point_cloud(100000)
    dev = [torch.from_numpy(a).cuda() for a in (P, G, R, AREA)]
    f = lambda: mh.point_pressure_batch(dev[0][None], dev[1], dev[2], lon, lat, AREA=dev[3], LMAX=60, LOVE=LOVE, as_numpy=False)
    med, best = timeit(f, n=10)
    print(json.dumps({'cfg': 'c2 1e5 points L60', 'ms_med': med, 'ms_best': best, 'tflops_alg': 2.18e9 / best / 1e9}))
if 'c4' in which:
    lon, lat = syn.grid_axes(721, 1440)
    G, R = syn.surface_geometry(lon, lat)
    P = syn.pressure_fields(721, 1440, 1)
    Pd, Gd, Rd = [torch.from_numpy(a).cuda() for a in (P, G, R)]
    L3 = syn.love_numbers(360)
    f = lambda: mh.pressure_stokes_batch(Pd, Gd, Rd, lon, lat, LMAX=360, LOVE=L3, as_numpy=False)
    med, best = timeit(f, n=5, warm=2)
    print(json.dumps({'cfg': 'c4 721x1440 L360', 'ms_med': med, 'ms_best': best, 'tflops_alg': 272e9 / best / 1e9}))
    f = lambda: mh.pressure_stokes_batch(Pd, Gd, Rd, lon, lat, LMAX=60, LOVE=LOVE, as_numpy=False)
    med, best = timeit(f, n=10)
    print(json.dumps({'cfg': '0.25deg L60 pressure', 'ms_med': med, 'ms_best': best, 'tflops_alg': 7.93e9 / best / 1e9}))
