import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import model_harmonics_b200 as mh
from model_harmonics_b200 import operators, testing as syn
lon, lat = syn.grid_axes(721, 1440)
GPH, dP, geoid = syn.atmosphere_cube(37, 721, 1440)
LOVE = syn.love_numbers(60)
kw = dict(LMAX=60, ELLIPSOID='WGS84', GEOID=geoid, LOVE=LOVE)
def pinned(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a).dtype, pin_memory=True); t.copy_(torch.from_numpy(a)); return t.numpy()
for dt in (np.float64, np.float32):
    g, d = pinned(GPH.astype(dt)), pinned(dP.astype(dt))
    for bands in (8, 1, 4, 16):
        operators._HOST_BANDS = bands
        Y = mh.gen_atmosphere_stokes(g, d, lon, lat, **kw)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5): Y = mh.gen_atmosphere_stokes(g, d, lon, lat, **kw)
        torch.cuda.synchronize(); dtm = (time.perf_counter() - t0) / 5
        print(dt.__name__, 'bands', bands, 'ms/field %.2f  fields/s %.1f  GB/s %.1f' % (1e3 * dtm, 1 / dtm, g.nbytes * 2 / dtm / 1e9))
    # device-path reference result
    Yd = mh.atmosphere_stokes_batch(torch.from_numpy(g).cuda()[None], torch.from_numpy(d).cuda()[None], lon, lat, **kw)
    print('  host vs device path rel diff', np.abs(Y.clm - Yd[0][0]).max() / np.abs(Y.clm).max())
