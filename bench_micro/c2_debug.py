import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import model_harmonics_b200 as mh
from model_harmonics_b200 import testing as syn
LOVE = syn.love_numbers(60)
pts = syn.point_cloud(100000)
dv = [torch.from_numpy(a).cuda() for a in (pts[0], pts[1], pts[2], pts[5])]
f = lambda: mh.point_pressure_batch(dv[0][None], dv[1], dv[2], pts[3], pts[4], AREA=dv[3], LMAX=60, LOVE=LOVE, as_numpy=False)
for i in range(12):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); f(); b.record(); torch.cuda.synchronize()
    print(i, 'event ms %.3f wall ms %.3f' % (a.elapsed_time(b), 1e3 * (time.perf_counter() - t0)))
# with and without pole points
for pole in (True, False):
    pts2 = syn.point_cloud(100000, pole_points=pole)
    dv2 = [torch.from_numpy(a).cuda() for a in (pts2[0], pts2[1], pts2[2], pts2[5])]
    g = lambda: mh.point_pressure_batch(dv2[0][None], dv2[1], dv2[2], pts2[3], pts2[4], AREA=dv2[3], LMAX=60, LOVE=LOVE, as_numpy=False)
    g(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g(); b.record(); torch.cuda.synchronize()
    print('pole_points', pole, a.elapsed_time(b))
