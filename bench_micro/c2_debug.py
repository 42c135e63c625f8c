import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import model_harmonics_b200 as mh
from model_harmonics_b200 import testing as syn, operators, _lib
LOVE = syn.love_numbers(60)
pts = syn.point_cloud(100000)
dv = [torch.from_numpy(a).cuda() for a in pts]
f = lambda: mh.point_pressure_batch(dv[0][None], dv[1], dv[2], dv[3], dv[4], AREA=dv[5], LMAX=60, LOVE=LOVE, as_numpy=False)
for i in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); f(); b.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print(i, 'event ms %.3f  host-call ms %.3f  total wall ms %.3f' % (a.elapsed_time(b), 1e3 * (t1 - t0), 1e3 * (time.perf_counter() - t0)))
# break down host side
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): f()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(12)
