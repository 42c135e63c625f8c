"""Runs one operator a few times (for ncu).  usage: prof_case.py c3|c1|c2|c4 [reps]"""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
import model_harmonics_b200 as mh
from model_harmonics_b200 import testing as syn
case = sys.argv[1]; reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
LOVE = syn.love_numbers(60)
if case == 'c3':
    lon, lat = syn.grid_axes(721, 1440)
    GPH, dP, geoid = syn.atmosphere_cube(37, 721, 1440)
    g, d, ge = torch.from_numpy(GPH).cuda(), torch.from_numpy(dP).cuda(), torch.from_numpy(geoid).cuda()
    f = lambda: mh.atmosphere_stokes_batch(g[None], d[None], lon, lat, as_numpy=False, LMAX=60, ELLIPSOID='WGS84', GEOID=ge, LOVE=LOVE)
elif case == 'c3b':
    # batch of 8 DISTINCT cubes (8 x 615 MB)
    lon, lat = syn.grid_axes(721, 1440)
    GPH, dP, geoid = syn.atmosphere_cube(37, 721, 1440)
    sc = np.array([syn.atmosphere_field_scalars(t) for t in range(8)])
    g = torch.from_numpy(GPH).cuda(); d = torch.from_numpy(dP).cuda(); ge = torch.from_numpy(geoid).cuda()
    g8 = (g[None] * torch.from_numpy(sc[:, 0]).cuda()[:, None, None, None]).contiguous()
    d8 = (d[None] * torch.from_numpy(sc[:, 1]).cuda()[:, None, None, None]).contiguous()
    del g, d
    f = lambda: mh.atmosphere_stokes_batch(g8, d8, lon, lat, as_numpy=False, LMAX=60, ELLIPSOID='WGS84', GEOID=ge, LOVE=LOVE)
elif case == 'c1':
    lon, lat = syn.grid_axes(181, 360); G, R = syn.surface_geometry(lon, lat); P = syn.pressure_fields(181, 360, 12)
    Pd, Gd, Rd = [torch.from_numpy(a).cuda() for a in (P, G, R)]
    f = lambda: mh.pressure_stokes_batch(Pd, Gd, Rd, lon, lat, LMAX=60, LOVE=LOVE, as_numpy=False)
elif case == 'c2':
    P, G, R, lon, lat, AREA = syn.point_cloud(100000)
    dev = [torch.from_numpy(a).cuda() for a in (P, G, R, AREA)]
    f = lambda: mh.point_pressure_batch(dev[0][None], dev[1], dev[2], lon, lat, AREA=dev[3], LMAX=60, LOVE=LOVE, as_numpy=False)
elif case == 'c4':
    lon, lat = syn.grid_axes(721, 1440); G, R = syn.surface_geometry(lon, lat); P = syn.pressure_fields(721, 1440, 1)
    Pd, Gd, Rd = [torch.from_numpy(a).cuda() for a in (P, G, R)]
    L3 = syn.love_numbers(360)
    f = lambda: mh.pressure_stokes_batch(Pd, Gd, Rd, lon, lat, LMAX=360, LOVE=L3, as_numpy=False)
for _ in range(reps):
    f()
torch.cuda.synchronize()
print('ok')
