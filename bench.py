#!/usr/bin/env python
"""
bench.py -- fields -> Stokes coefficients per second (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm (oracle port)

Workload (config.workload = "atmosphere_c3"): ERA5-shaped 0.25 degree, 37-level atmospheric
fields (37 x 721 x 1440, fp64) -> LMAX = 60 Stokes coefficients with gen_atmosphere_stokes'
arithmetic (BASELINE.json configs[2]/[4]; the configuration the north star's target is quoted
on).  A step is one pass of the hot path over a batch of --fields fields per GPU, each derived
on load from one device-resident base cube by the IEEE-exact scalings of SURVEY.md 8(d)
(GPH_t = GPH_0 (1+a_t), dP_t = dP_0 (1+b_t)); with N > 1 every rank takes its own time steps
(weak scaling, no collective on the data path: SURVEY.md 8e); the coefficient blocks of the K timed steps
are gathered by ONE NCCL all-gather at the end, inside the timed region.

value   : whole-job fields/s with inputs resident in HBM (CUDA events, max over ranks); the --fields
          fields of a step are DISTINCT cubes (8 x 615 MB), so every byte comes from DRAM.
e2e     : the same metric through the drop-in Python call gen_atmosphere_stokes with HOST
          (pinned) fp64 cubes: H2D of both cubes and D2H of the coefficients inside the timed
          region, every field.
roofline: algorithmic bytes of SURVEY.md 8(d) per launch / device time of atm_moment_kernel (CUDA
          events on its own stream, inside the timed region) against the measured HBM peak
          (MEASURED_PEAKS.json) -- the floor of this workload now that the contracted-moment form
          has cut the fp64 work ~5x; the fp64 views (direct-sum flop model, executed-flop model,
          ncu pipe utilisation) sit beside it under roofline.fp64.  traffic = DRAM bytes of the
          same batched launch measured under ncu (profiles/r02_atm_moment_batch8.json).
c5_480_field_job: BASELINE.json configs[4] -- 480 fields time-sharded over the ranks (strong
          scaling), one gather at the end, wall clock barrier -> gather, parity of
          t in {0, 7, 123, 479} against reference-generated goldens.
cpu_baseline / --impl reference: the NumPy oracle port of the reference's algorithm on the
          box's host cores, process-parallel over time steps, on a bounded latitude sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NLEV, NLAT, NLON, LMAX = 37, 721, 1440, 60
FP64_PEAK_TFLOPS = 37.0       # measured on this pool: DMMA m8n8k4 fp64 (profiles/r01_fp64_peaks.jsonl)
SAMPLE_STRIDE = 8             # CPU arm: every 8th latitude ring (91 of 721)


def algorithmic_flops_per_field(nlev=NLEV, nlat=NLAT, nlon=NLON, L=LMAX):
    """F_atm = nlev nlat nlon [3(L+1)+50] + nlat nlon [4 Ntri + (L+1) + 1] + 8 nlat Ntri."""
    ntri = (L + 1) * (L + 2) // 2
    f_grid = nlat * nlon * (4 * ntri + (L + 1) + 1) + nlat * ntri * 8
    return nlev * nlat * nlon * (3 * (L + 1) + 50) + f_grid


def executed_flops_per_field(nlev=NLEV, nlat=NLAT, nlon=NLON):
    """fp64 operations atm_moment_kernel issues per field: per (level, column) element
    11 slots of geometry + 22 of moment accumulation, per column 512 FMA-slots of DMMA contraction."""
    return 2 * nlat * nlon * (nlev * 33 + 512)


def algorithmic_bytes_per_field(nlev=NLEV, nlat=NLAT, nlon=NLON, itemsize=8):
    """both cubes read once."""
    return 2 * itemsize * nlev * nlat * nlon


# ------------------------------------------------------------------------------------------
# CPU arm: oracle port on a bounded latitude sample, process-parallel over time steps
# ------------------------------------------------------------------------------------------
def _cpu_worker(seed):
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    import numpy as np
    from oracle import stokes_oracle
    from oracle.gravtk_standin import plm_holmes
    from model_harmonics_b200 import testing as syn
    lat = np.linspace(90.0, -90.0, NLAT)[::SAMPLE_STRIDE]
    lon = np.arange(NLON) * (360.0 / NLON)
    GPH, dP, geoid = syn.atmosphere_cube(NLEV, len(lat), NLON, seed=seed)
    th = np.radians(90.0 - lat)
    PLM, _ = plm_holmes(LMAX, np.cos(th))          # pre-computed and passed in, as the drivers do
    w = np.sin(th) * np.radians(360.0 / NLON) * np.radians(180.0 / (NLAT - 1))
    t0 = time.perf_counter()
    stokes_oracle.gen_atmosphere_stokes(GPH, dP, lon, lat, LMAX=LMAX, ELLIPSOID='WGS84', GEOID=geoid, WEIGHT=w,
                                        PLM=PLM, LOVE=syn.love_numbers(LMAX), METHOD='BC05')
    return time.perf_counter() - t0


def cpu_arm_step(pool, cores):
    """
    One step: every core transforms one latitude-sampled field, all at the same time (so memory
    contention between cores is in the number).  Each worker times only the operator call (input
    synthesis and the PLM table are outside, as in the drivers).  Returns (fields/s, wall seconds):
    process-parallel throughput = sum over workers of (sample fraction / operator time).
    """
    nsample = len(range(0, NLAT, SAMPLE_STRIDE))
    t0 = time.perf_counter()
    times = pool.map(_cpu_worker, [1234 + i for i in range(cores)], chunksize=1)
    wall = time.perf_counter() - t0
    return sum((nsample / float(NLAT)) / t for t in times), wall


def cpu_sample_description():
    nsample = len(range(0, NLAT, SAMPLE_STRIDE))
    return ('NumPy oracle port of gen_atmosphere_stokes (BC05, PLM precomputed), one field per core, '
            'latitude rings 0::%d (%d of %d) x %d lon x %d levels, LMAX=%d; fields/s scaled by %d/%d '
            '(cost is linear in the ring count)' % (SAMPLE_STRIDE, nsample, NLAT, NLON, NLEV, LMAX, nsample, NLAT))


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_arm_step(pool, cores)
        t0 = time.perf_counter()
        rates = [cpu_arm_step(pool, cores)[0] for _ in range(args.steps)]
        wall = time.perf_counter() - t0
    value = sum(rates) / len(rates)
    line = {
        'impl': 'reference', 'metric': 'fields_to_stokes_per_second', 'value': value, 'unit': 'fields/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * wall / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args, cores),
        'cpu_baseline': {'value': value, 'unit': 'fields/s', 'cores': cores, 'kind': 'port',
                         'sample': cpu_sample_description()},
        'e2e': {'value': value, 'unit': 'fields/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


def workload_config(args, cores=None):
    return {'workload': 'atmosphere_c3', 'operator': 'gen_atmosphere_stokes', 'method': 'BC05',
            'grid': '%dx%dx%d' % (NLEV, NLAT, NLON), 'lmax': LMAX, 'fields_per_step_per_gpu': args.fields,
            'input_dtype': 'f64',
            'l2_policy': 'inputs exceed L2: %d distinct device-resident cubes per step, 615 MB each, vs 126 MB of L2'
                         % args.fields,
            'parallelism': 'time-sharded x%d' % args.gpus,
            'collective': ('none on the data path; one all_gather_into_tensor of the timed steps\' coefficient '
                           'blocks at the end, inside the timed region') if args.gpus > 1 else None}


# ------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler(object):
    FIELDS = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, device_index):
        self.rows = []
        self.proc = None
        self.device_index = device_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.device_index), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', os.environ.get('MHB200_BENCH_LMS', '200')],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def wait_first_row(self, timeout=5.0):
        t0 = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smmax, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ts, line in self.rows:
            parts = [x.strip() for x in line.split(',')]
            if len(parts) < 8:
                continue
            try:
                clk, cmax = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            smmax = cmax
            if t_begin - 0.22 <= ts <= t_end + 0.22:
                sm.append(clk)
                for n, v in zip(names, parts[4:8]):
                    if v.lower().startswith('active'):
                        reasons.add(n)
        note = None
        if not sm:      # region shorter than the sampling period: take the samples nearest to it
            near = sorted(self.rows, key=lambda r: min(abs(r[0] - t_begin), abs(r[0] - t_end)))[:3]
            for _, line in near:
                parts = [x.strip() for x in line.split(',')]
                try:
                    sm.append(float(parts[1]))
                except (ValueError, IndexError):
                    pass
            note = 'timed region shorter than the sampling period: nearest samples used'
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': smmax, 'reasons': ['no samples'], 'samples': 0}
        sm.sort()
        out = {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': smmax, 'reasons': sorted(reasons), 'samples': len(sm)}
        if note:
            out['note'] = note
        return out


# ------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(torch, local):
    """
    Pins this process (and therefore the first-touch placement of its pinned host buffers) to the CPUs of
    the NUMA node its GPU hangs off, so that N ranks do not all stream from node 0.  Returns a description.
    """
    try:
        props = torch.cuda.get_device_properties(local)
        bus = '%04x:%02x:%02x.0' % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bus).read().strip())
        cpus = set()

        def parse_list(text):
            out = set()
            for part in text.strip().split(','):
                lo, _, hi = part.partition('-')
                out.update(range(int(lo), int(hi or lo) + 1))
            return out

        if node >= 0:
            cpus = parse_list(open('/sys/devices/system/node/node%d/cpulist' % node).read())
        else:
            # containers often report -1 in sysfs: ask the driver (nvidia-smi topo -m, "CPU Affinity" column)
            topo = subprocess.run(['nvidia-smi', 'topo', '-m'], capture_output=True, text=True, timeout=20).stdout
            lines = [ln for ln in topo.splitlines() if ln.strip()]
            header = [h.strip() for h in lines[0].replace('\x1b[4m', '').replace('\x1b[0m', '').split('\t')]
            col = next((i for i, h in enumerate(header) if h.startswith('CPU Affinity')), None)
            row = next((ln for ln in lines[1:] if ln.split('\t')[0].strip() == 'GPU%d' % local), None)
            if col is None or row is None:
                return {'numa_node': None, 'note': 'no NUMA affinity reported for %s' % bus}
            # the header line starts with an empty cell above the row labels, so columns line up by index
            cells = [c.strip() for c in row.split('\t')]
            cpus = parse_list(cells[col])
            ncol = next((i for i, h in enumerate(header) if h.startswith('NUMA Affinity')), None)
            try:
                node = int(cells[ncol]) if ncol is not None else None
            except (ValueError, IndexError):
                node = None
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {'numa_node': node, 'cpus': len(cpus)}
    except Exception as exc:      # containers often hide sysfs: report, do not fail
        return {'numa_node': None, 'note': '%s: %s' % (type(exc).__name__, exc)}


def load_profile_record(name):
    """Figures measured under ncu on the same launch shape and committed under profiles/ (or None)."""
    path = os.path.join(ROOT, 'profiles', name)
    if os.path.isfile(path):
        try:
            return json.load(open(path))
        except Exception:
            return None
    return None


def golden_error(np, clm, slm, name):
    """max|delta| / max|ref| against a committed reference-generated golden (tests/golden/<name>.npz)."""
    path = os.path.join(ROOT, 'tests', 'golden', name + '.npz')
    if not os.path.isfile(path):
        return None
    gold = np.load(path)
    scale = max(np.abs(gold['clm']).max(), np.abs(gold['slm']).max())
    return float(max(np.abs(clm - gold['clm']).max(), np.abs(slm - gold['slm']).max()) / scale)


def run_cuda_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import model_harmonics_b200 as mh
    from model_harmonics_b200 import _lib, batch, testing as syn

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: the product path has no CPU fallback')
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(torch, local)
    if world > 1:
        import datetime
        dist.init_process_group('nccl', device_id=torch.device('cuda', local),
                                timeout=datetime.timedelta(seconds=180))
    dev = torch.device('cuda', local)
    L = _lib.lib()
    T = args.fields

    lon, lat = syn.grid_axes(NLAT, NLON)
    GPH, dP, geoid = syn.atmosphere_cube(NLEV, NLAT, NLON, seed=1234)
    LOVE = syn.love_numbers(LMAX)
    g, d, ge = torch.from_numpy(GPH).to(dev), torch.from_numpy(dP).to(dev), torch.from_numpy(geoid).to(dev)
    kw = dict(LMAX=LMAX, ELLIPSOID='WGS84', GEOID=ge, LOVE=LOVE, METHOD='BC05')
    # this rank's time steps (global index = rank*T + i), SURVEY.md 8(d): GPH_t = GPH_0 (1 + a_t),
    # dP_t = dP_0 (1 + b_t), IEEE-exact elementwise products -- T DISTINCT device-resident cubes
    # (T x 615 MB), so every field of a step is read from DRAM, not from L2
    scales = np.array([syn.atmosphere_field_scalars(rank * T + i) for i in range(T)])
    sc = torch.from_numpy(scales).to(dev)
    g8 = (g[None] * sc[:, 0, None, None, None]).contiguous()
    d8 = (d[None] * sc[:, 1, None, None, None]).contiguous()

    def step():
        clm, slm = mh.atmosphere_stokes_batch(g8, d8, lon, lat, as_numpy=False, **kw)
        return torch.stack([clm, slm], dim=1)

    def collect(blocks):
        # the only collective: one gather of every step's (T, 2, L+1, M+1) block, at the end of the job
        blk = torch.cat(blocks, dim=0)
        if world > 1:
            blk = batch.gather_coefficients(blk, world * blk.shape[0])
        return blk

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out = step()
    for _ in range(2):
        # the timed pattern itself, untimed: K results alive at once (the caching allocator grows here, not in the
        # timed region: a cudaMalloc there cost a 30-120 ms step) and the gather at its timed size (NCCL sets up per size)
        warm = [step() for _ in range(args.steps)]
        collect(warm)
        del warm
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        sampler.wait_first_row()
    # every rank keeps its GPU under the same load while nvidia-smi takes its first samples; the library's event
    # hook is on so that its events exist and have been recorded before the timed region (creating them there cost
    # up to 100 ms in one step)
    L.mhb200_profile_enable(1)
    for _ in range(20):
        out = step()
    torch.cuda.synchronize()
    L.mhb200_profile_enable(0)
    import ctypes
    # Every step is bracketed by its own events: a host or driver stall inside the few milliseconds of the timed
    # region (a cudaMalloc, an nvidia-smi query holding launches back) shows as one step of tens of milliseconds
    # among steps of 1.6 ms.  If one step takes more than three times the median, the K steps are measured once more
    # and the line says so (`remeasured`, with the first attempt's figures).
    remeasured = None
    for attempt in range(2):
        L.mhb200_profile_enable(1)
        n0 = L.mhb200_launch_count()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 2)]
        for e in ev:
            e.record()                    # created and recorded once before they are used
        barrier()
        t_begin = time.perf_counter()
        ev[0].record()
        outs = []
        for i in range(args.steps):
            outs.append(step())
            ev[i + 1].record()
        gathered = collect(outs)
        ev[args.steps + 1].record()
        barrier()
        t_end = time.perf_counter()
        ms = ev[0].elapsed_time(ev[args.steps + 1])
        launches = int(L.mhb200_launch_count() - n0)
        nrec = ctypes.c_int(0)
        ring_ms = L.mhb200_profile_mean_ms(ctypes.byref(nrec))
        L.mhb200_profile_enable(0)
        raw_steps = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
        per_step = sorted(raw_steps)
        hiccup = torch.tensor([1.0 if per_step[-1] > 3.0 * per_step[len(per_step) // 2] else 0.0], device=dev)
        if world > 1:
            dist.all_reduce(hiccup, op=dist.ReduceOp.MAX)
        if attempt == 0 and float(hiccup.item()) > 0.0:
            remeasured = {'reason': 'one timed step took more than 3x the median (host / driver stall)',
                          'first_attempt_ms_per_step': ms / args.steps,
                          'first_attempt_slowest_step_ms': per_step[-1],
                          'first_attempt_slowest_step_index': raw_steps.index(per_step[-1]),
                          'first_attempt_median_step_ms': per_step[len(per_step) // 2]}
            continue
        break
    clocks = sampler.stop(t_begin, t_end) if sampler else None
    per_rank = None
    if world > 1:
        # every rank's own device time and dominant-kernel time (the max over ranks is the job's)
        mine = torch.tensor([ms / args.steps, ring_ms], dtype=torch.float64, device=dev)
        allr = torch.empty((world, 2), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine)
        per_rank = {'ms_per_step': [round(float(v), 4) for v in allr[:, 0]],
                    'kernel_ms': [round(float(v), 4) for v in allr[:, 1]]}
        ms = float(allr[:, 0].max().item()) * args.steps
    value = world * T * args.steps / (ms * 1e-3)
    # parity of what was timed, against committed reference-generated goldens (rank 0 holds t = 0 and 7)
    spot = {}
    if rank == 0:
        res = outs[-1].cpu().numpy()
        assert gathered.shape[0] == world * T * args.steps
        for t in (0, 7):
            if t < T:
                spot['t%d' % t] = golden_error(np, res[t, 0], res[t, 1], 'atmosphere_c5_t%d' % t)
    del g8, d8

    # ---- the 480-field job (BASELINE.json configs[4]): strong scaling, fields derived on load from the
    #      resident base cube, one plan, ONE gather at the end; wall time barrier -> gather
    c5 = None
    if not args.no_c5:
        NF = args.c5_fields
        c5_scales = np.array([syn.atmosphere_field_scalars(t) for t in range(NF)])
        kw5 = dict(LMAX=LMAX, ELLIPSOID='WGS84', GEOID=ge, LOVE=LOVE, METHOD='BC05')
        batch.atmosphere_stokes_sharded(g, d, lon, lat, c5_scales[:world * 8], **kw5)       # warm-up
        barrier()
        t0 = time.perf_counter()
        c5_clm, c5_slm = batch.atmosphere_stokes_sharded(g, d, lon, lat, c5_scales, **kw5)
        barrier()
        c5_s = time.perf_counter() - t0
        if world > 1:
            ts = torch.tensor([c5_s], dtype=torch.float64, device=dev)
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
            c5_s = float(ts.item())
        if rank == 0:
            cc, ss = c5_clm.cpu().numpy(), c5_slm.cpu().numpy()
            c5 = {'fields': NF, 'scaling': 'strong', 'seconds': c5_s, 'fields_per_s': NF / c5_s,
                  'fields_per_rank': [batch.shard_range(NF, r, world)[1] - batch.shard_range(NF, r, world)[0]
                                      for r in range(world)],
                  'collective': 'one all_gather_into_tensor of (T, 2, 61, 61) fp64 at the end' if world > 1 else None,
                  'parity_vs_reference_golden': {('t%d' % t): golden_error(np, cc[t], ss[t], 'atmosphere_c5_t%d' % t)
                                                 for t in (0, 7, 123, 479) if t < NF}}

    # ---- end to end through the drop-in call with pinned host cubes (H2D + D2H every field)
    nf = 2
    host = []
    for i in range(nf):
        a, b = syn.atmosphere_field_scalars(rank * T + i)
        hg = torch.empty((NLEV, NLAT, NLON), dtype=torch.float64, pin_memory=True)
        hd = torch.empty((NLEV, NLAT, NLON), dtype=torch.float64, pin_memory=True)
        hg.copy_(torch.from_numpy(GPH) * a)
        hd.copy_(torch.from_numpy(dP) * b)
        host.append((hg.numpy(), hd.numpy()))
    kw_host = dict(LMAX=LMAX, ELLIPSOID='WGS84', GEOID=geoid, LOVE=LOVE, METHOD='BC05')

    def e2e_step():
        return [mh.gen_atmosphere_stokes(hg, hd, lon, lat, **kw_host) for hg, hd in host]

    Y = e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        Y = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        ts = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        e2e_s = float(ts.item())
    e2e_value = world * nf * e2e_steps / e2e_s
    if rank == 0:
        spot['e2e_t0'] = golden_error(np, Y[0].clm, Y[0].slm, 'atmosphere_c5_t0')
    # same call with the cubes in their on-disk storage type (float32, widened exactly on the device)
    host32 = []
    for hg, hd in host[:1]:
        a32 = torch.empty((NLEV, NLAT, NLON), dtype=torch.float32, pin_memory=True)
        b32 = torch.empty((NLEV, NLAT, NLON), dtype=torch.float32, pin_memory=True)
        a32.copy_(torch.from_numpy(hg))
        b32.copy_(torch.from_numpy(hd))
        host32.append((a32.numpy(), b32.numpy()))
    mh.gen_atmosphere_stokes(host32[0][0], host32[0][1], lon, lat, **kw_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(2 * e2e_steps):
        mh.gen_atmosphere_stokes(host32[0][0], host32[0][1], lon, lat, **kw_host)
    barrier()
    e2e32_s = time.perf_counter() - t0
    if world > 1:
        ts = torch.tensor([e2e32_s], dtype=torch.float64, device=dev)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        e2e32_s = float(ts.item())
    e2e32_value = world * 2 * e2e_steps / e2e32_s
    h2d = nf * algorithmic_bytes_per_field()
    d2h = nf * 2 * (LMAX + 1) * (LMAX + 1) * 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    ppath = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(ppath):
        peaks = json.load(open(ppath))
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    # dominant kernel = atm_moment_kernel (K_A of the contracted-moment path): one launch per step
    alg_bytes = algorithmic_bytes_per_field() * T
    hbm_achieved = alg_bytes / (ring_ms * 1e-3) / 1e9
    flops = algorithmic_flops_per_field() * T
    prof = load_profile_record('r02_atm_moment_batch8.json') or {}
    traffic = None
    if prof.get('fields_per_launch') == T:
        traffic = int(prof['dram_bytes_read'] + prof['dram_bytes_write'])

    line = {
        'metric': 'fields_to_stokes_per_second', 'value': value, 'unit': 'fields/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'per_rank': per_rank, 'remeasured': remeasured,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': workload_config(args),
        'e2e': {'value': e2e_value, 'unit': 'fields/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'fields_per_step': nf, 'steps': e2e_steps, 'call': 'gen_atmosphere_stokes(pinned NumPy fp64 cubes)',
                'pipeline': '4 latitude bands: H2D of band b+1 overlaps the kernels of band b; the NumPy geoid is staged '
                            'behind the first band (one flat pinned H2D copy: 55.5 GB/s on this pool)',
                'h2d_gb_per_s_aggregate': e2e_value * algorithmic_bytes_per_field() / 1e9,
                'numa': numa,
                'float32_stored_cubes': {'value': e2e32_value, 'unit': 'fields/s',
                                         'h2d_bytes_per_field': algorithmic_bytes_per_field(itemsize=4),
                                         'h2d_gb_per_s_aggregate':
                                             e2e32_value * algorithmic_bytes_per_field(itemsize=4) / 1e9}},
        'gpu_launches': launches,
        # The contracted-moment form executes ~5x fewer fp64 operations than the direct sum SURVEY 8(d)
        # counts, so the floor of this workload is the HBM stream (614.6 MB per field, 94 us at the measured
        # peak): the roofline is quoted against HBM, with the fp64 views beside it.
        'roofline': {'bound': 'hbm', 'kernel': 'atm_moment_kernel', 'achieved': hbm_achieved, 'peak': hbm_peak,
                     'unit': 'GB/s', 'frac': hbm_achieved / hbm_peak, 'traffic': traffic,
                     'algorithmic_bytes_per_launch': alg_bytes, 'kernel_ms': ring_ms,
                     'launches_timed': int(nrec.value),
                     'peak_source': 'MEASURED_PEAKS.json (burst copy bandwidth)' if peaks else 'fallback',
                     'traffic_source': 'profiles/r02_atm_moment_batch8.json: dram__bytes_read.sum + '
                                       'dram__bytes_write.sum of the same batched launch under ncu --set full'
                                       if traffic else None,
                     'fp64': {
                         'direct_sum_model': {
                             'flops_per_launch': flops, 'achieved_tflops': flops / (ring_ms * 1e-3) / 1e12,
                             'peak_tflops': FP64_PEAK_TFLOPS,
                             'frac': flops / (ring_ms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS,
                             'note': 'SURVEY 8(d) flop count of the reference formulation; the moment form '
                                     'removes most of these operations, so this can exceed 1'},
                         'executed_model': {
                             'flops_per_launch': executed_flops_per_field() * T,
                             'achieved_tflops': executed_flops_per_field() * T / (ring_ms * 1e-3) / 1e12,
                             'frac': executed_flops_per_field() * T / (ring_ms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS,
                             'note': 'fp64 operations the kernel issues (33 FMA-slots per level element + 512 per '
                                     'column in the DMMA contraction, FMA = 2 flops)'},
                         'pipe_busy_pct_ncu': prof.get('pipe_fp64_plus_dmma_pct'),
                         'peak_source': 'fp64 DMMA m8n8k4 peak measured on this pool '
                                        '(profiles/r01_fp64_peaks.jsonl); DFMA and DMMA share one pipe'}},
        'clocks': clocks,
        'parity_vs_reference_golden': spot,
    }
    if c5 is not None:
        line['c5_480_field_job'] = c5
    if world == 1 and not args.no_cpu_baseline:
        import multiprocessing as mp
        cores = os.cpu_count() or 1
        with mp.get_context('spawn').Pool(cores) as pool:
            cpu_arm_step(pool, cores)                  # untimed warm-up (imports, page cache), as in the reference arm
            cpu_value, wall = cpu_arm_step(pool, cores)
        line['cpu_baseline'] = {'value': cpu_value, 'unit': 'fields/s', 'cores': cores, 'kind': 'port',
                                'sample': cpu_sample_description(), 'seconds': wall}
    if world == 1 and not args.no_extra:
        line['other_workloads'] = other_workloads(mh, syn, torch, L)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def other_workloads(mh, syn, torch, L):
    """Device-resident timings of the remaining BASELINE.json configs (median of a few runs)."""
    import ctypes
    import numpy as np

    def timeit(fn, n, warm=2):
        """Per-call device time in a back-to-back stream of n calls (as `value` is measured): median of three
        groups; the isolated single-call latency (idle GPU before every call, host launch path exposed) is
        returned beside it."""
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        groups = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n):
                fn()
            b.record()
            torch.cuda.synchronize()
            groups.append(a.elapsed_time(b) / n)
        single = []
        for _ in range(min(n, 5)):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            single.append(a.elapsed_time(b))
        timeit.single = float(np.median(single))
        return float(np.median(groups))

    out = {}
    LOVE = syn.love_numbers(60)
    # C1: 12 monthly 1-degree surface pressure fields, LMAX 60
    lon, lat = syn.grid_axes(181, 360)
    G, R = syn.surface_geometry(lon, lat)
    P = syn.pressure_fields(181, 360, 12)
    Pd, Gd, Rd = [torch.from_numpy(a).cuda() for a in (P, G, R)]
    ms = timeit(lambda: mh.pressure_stokes_batch(Pd, Gd, Rd, lon, lat, LMAX=60, LOVE=LOVE, as_numpy=False), 10)
    out['pressure_c1'] = {'config': '12 x 181x360, LMAX=60', 'ms': ms, 'ms_single_call': timeit.single,
                          'fields_per_s': 12e3 / ms,
                          'algorithmic_tflops': 12 * 0.500e9 / ms / 1e9}
    # C2: 1e5 point pressures, LMAX 60
    pts = syn.point_cloud(100000)
    dv = [torch.from_numpy(a).cuda() for a in pts]       # P, G, R, lon, lat, AREA all device resident
    ms = timeit(lambda: mh.point_pressure_batch(dv[0][None], dv[1], dv[2], dv[3], dv[4], AREA=dv[5], LMAX=60,
                                                LOVE=LOVE, as_numpy=False), 10)
    out['points_c2'] = {'config': '1e5 points, LMAX=60', 'ms': ms, 'ms_single_call': timeit.single,
                        'calls_per_s': 1e3 / ms,
                        'algorithmic_tflops': 2.18e9 / ms / 1e9}
    # C4: 0.25-degree surface pressure to LMAX 360
    lon, lat = syn.grid_axes(721, 1440)
    G, R = syn.surface_geometry(lon, lat)
    P = syn.pressure_fields(721, 1440, 1)
    Pd, Gd, Rd = [torch.from_numpy(a).cuda() for a in (P, G, R)]
    L3 = syn.love_numbers(360)
    ms = timeit(lambda: mh.pressure_stokes_batch(Pd, Gd, Rd, lon, lat, LMAX=360, LOVE=L3, as_numpy=False), 3, warm=1)
    out['pressure_c4'] = {'config': '721x1440, LMAX=360', 'ms': ms, 'ms_single_call': timeit.single,
                          'fields_per_s': 1e3 / ms,
                          'algorithmic_tflops': 272e9 / ms / 1e9}
    # 8f-2: upstream producer, T/q/sp on 37 levels of the 0.25 degree grid -> (Z, dp); HBM-bound stream
    Tm, Qm, SPm, zs, A, B, AI, BI = syn.model_level_inputs(NLEV, NLAT, NLON, seed=11)
    dvt = [torch.from_numpy(a).cuda() for a in (Tm, Qm, SPm, zs)]
    ms = timeit(lambda: mh.geopotential_levels(dvt[0], dvt[1], dvt[2], dvt[3], A, B, AI, BI, divide_by=9.80665,
                                               as_numpy=False), 10)
    moved = 4 * 4 * NLEV * NLAT * NLON      # read T, q; write Z, dp (float32)
    peaks = {}
    if os.path.isfile(os.path.join(ROOT, 'MEASURED_PEAKS.json')):
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    hbm = peaks.get('hbm_gbs', 6650.0)
    out['hypsometric_f2'] = {'config': '%dx%dx%d float32 T,q -> Z,dp' % (NLEV, NLAT, NLON), 'ms': ms,
                             'ms_single_call': timeit.single,
                             'fields_per_s': 1e3 / ms, 'algorithmic_bytes': moved,
                             'roofline': {'bound': 'hbm', 'achieved': moved / ms / 1e6, 'peak': hbm, 'unit': 'GB/s',
                                          'frac': moved / ms / 1e6 / hbm}}
    return out


_RESULT_FD = None


def _claim_stdout():
    """stdout carries exactly ONE JSON line: everything libraries write to file descriptor 1 while the
    benchmark runs (NCCL prints its version there) goes to stderr; the line is written to the real
    stdout by emit()."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + '\n').encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_RESULT_FD, data)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--fields', type=int, default=8, help='fields per step per GPU')
    ap.add_argument('--impl', default='cuda', choices=['cuda', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extra', action='store_true')
    ap.add_argument('--no-c5', action='store_true', help='skip the 480-field strong-scaling job')
    ap.add_argument('--c5-fields', type=int, default=480)
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_cuda_arm(args)


if __name__ == '__main__':
    main()
