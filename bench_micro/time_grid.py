"""Device-resident timings of the grid operators (CUDA events), for A/B runs on the GPU box.
usage: python bench_micro/time_grid.py [c3] [c3batch] [c4] [c1]   (env MHB200_MOMENTS etc. apply)"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import model_harmonics_b200 as mh
from model_harmonics_b200 import testing as syn

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    return ms[len(ms) // 2], ms[0]


def rel(c, s, gold):
    sc = max(np.abs(gold['clm']).max(), np.abs(gold['slm']).max())
    return float(max(np.abs(c - gold['clm']).max(), np.abs(s - gold['slm']).max()) / sc)


def main():
    what = sys.argv[1:] or ['c3', 'c3batch', 'c4', 'c1']
    out = {'MHB200_MOMENTS': os.environ.get('MHB200_MOMENTS', '')}
    L = 60
    LOVE = syn.love_numbers(L)
    if 'c3' in what or 'c3batch' in what:
        lon, lat = syn.grid_axes(721, 1440)
        GPH, dP, geoid = syn.atmosphere_cube(37, 721, 1440)
        g = torch.from_numpy(GPH).cuda()
        d = torch.from_numpy(dP).cuda()
        kw = dict(LMAX=L, ELLIPSOID='WGS84', GEOID=torch.from_numpy(geoid).cuda(), LOVE=LOVE, as_numpy=False)
        for method in ('BC05', 'SW02'):
            c, s = mh.atmosphere_stokes_batch(g[None], d[None], lon, lat, METHOD=method, **kw)
            gold = np.load(os.path.join(GOLDEN, 'atmosphere_c3_%s.npz' % method.lower()))
            out['c3_%s_rel' % method] = rel(c[0].cpu().numpy(), s[0].cpu().numpy(), gold)
            med, best = timed(lambda: mh.atmosphere_stokes_batch(g[None], d[None], lon, lat, METHOD=method, **kw))
            out['c3_%s_ms' % method] = (med, best)
        if 'c3batch' in what:
            T = 8
            scales = np.array([syn.atmosphere_field_scalars(t) for t in range(T)])
            med, best = timed(lambda: mh.atmosphere_stokes_batch(g, d, lon, lat, field_scale=scales, **kw), reps=6)
            out['c3_shared_cube_x8_ms_per_field'] = (med / T, best / T)
            g8 = (g[None] * torch.from_numpy(scales[:, 0]).cuda()[:, None, None, None]).contiguous()
            d8 = (d[None] * torch.from_numpy(scales[:, 1]).cuda()[:, None, None, None]).contiguous()
            c, s = mh.atmosphere_stokes_batch(g8, d8, lon, lat, **kw)
            gold = np.load(os.path.join(GOLDEN, 'atmosphere_c5_t7.npz'))
            out['c3_distinct_t7_rel'] = rel(c[7].cpu().numpy(), s[7].cpu().numpy(), gold)
            med, best = timed(lambda: mh.atmosphere_stokes_batch(g8, d8, lon, lat, **kw), reps=6)
            out['c3_distinct_cubes_x8_ms_per_field'] = (med / T, best / T)
            del g8, d8
        # float32-stored cubes
        g32, d32 = g.float(), d.float()
        med, best = timed(lambda: mh.atmosphere_stokes_batch(g32[None], d32[None], lon, lat, **kw))
        out['c3_f32_ms'] = (med, best)
        del g, d, g32, d32
    if 'c4' in what:
        lon, lat = syn.grid_axes(721, 1440)
        G, R = syn.surface_geometry(lon, lat)
        P = syn.pressure_fields(721, 1440, 1)
        Pd, Gd, Rd = (torch.from_numpy(a).cuda() for a in (P, G, R))
        for LL in (360, 60):
            LV = syn.love_numbers(LL)
            fn = lambda: mh.pressure_stokes_batch(Pd, Gd, Rd, lon, lat, LMAX=LL, LOVE=LV, as_numpy=False)
            c, s = fn()
            if LL == 360 and os.path.exists(os.path.join(GOLDEN, 'pressure_c4.npz')):
                out['c4_rel'] = rel(c[0].cpu().numpy(), s[0].cpu().numpy(), np.load(os.path.join(GOLDEN, 'pressure_c4.npz')))
            out['pressure_721x1440_L%d_ms' % LL] = timed(fn, reps=6)
    if 'c1' in what:
        lon, lat = syn.grid_axes(181, 360)
        G, R = syn.surface_geometry(lon, lat)
        P = syn.pressure_fields(181, 360, 12)
        Pd, Gd, Rd = (torch.from_numpy(a).cuda() for a in (P, G, R))
        fn = lambda: mh.pressure_stokes_batch(Pd, Gd, Rd, lon, lat, LMAX=60, LOVE=LOVE, as_numpy=False)
        c, s = fn()
        gold = np.load(os.path.join(GOLDEN, 'pressure_c1.npz'))
        out['c1_rel_t0'] = rel(c[0].cpu().numpy(), s[0].cpu().numpy(), gold)
        out['c1_12fields_ms'] = timed(fn)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
