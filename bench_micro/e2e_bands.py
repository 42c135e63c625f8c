"""e2e fields/s of gen_atmosphere_stokes(pinned host fp64 cubes) for the band count in MHB200_HOST_BANDS."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import model_harmonics_b200 as mh
from model_harmonics_b200 import testing as syn
lon, lat = syn.grid_axes(721, 1440)
GPH, dP, geoid = syn.atmosphere_cube(37, 721, 1440)
hg = torch.empty((37, 721, 1440), dtype=torch.float64, pin_memory=True); hg.copy_(torch.from_numpy(GPH))
hd = torch.empty((37, 721, 1440), dtype=torch.float64, pin_memory=True); hd.copy_(torch.from_numpy(dP))
kw = dict(LMAX=60, ELLIPSOID='WGS84', GEOID=(torch.from_numpy(geoid).cuda() if os.environ.get('GEOID_DEV') else geoid), LOVE=syn.love_numbers(60))
g, d = hg.numpy(), hd.numpy()
for _ in range(2): mh.gen_atmosphere_stokes(g, d, lon, lat, **kw)
torch.cuda.synchronize(); t0 = time.perf_counter()
n = 10
for _ in range(n): mh.gen_atmosphere_stokes(g, d, lon, lat, **kw)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print('bands', os.environ.get('MHB200_HOST_BANDS', '4'), 'geoid_dev', bool(os.environ.get('GEOID_DEV')), 'fields/s %.2f' % (n / dt), 'H2D GB/s %.1f' % (n * 614.6e6 / dt / 1e9))
# flat pinned copy of the same bytes for comparison
dev = torch.empty((37, 721, 1440), dtype=torch.float64, device='cuda')
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(n): dev.copy_(hg, non_blocking=True); dev.copy_(hd, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print('flat pinned H2D GB/s %.1f' % (n * 614.6e6 / dt / 1e9))
