#!/usr/bin/env python
"""bench.py — columns/sec of the xcape column hot path on B200 (headline: BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on host cores
    python bench.py --workload {C2,C3,C4,C5} ...             # C2 (default) is the headline
    (N > 1: launched by torchrun, one rank per GPU)

One "step" = one pass of the hot path over one synthetic field.  The default workload is
BASELINE configs[1] ("C2"): calc_cape, most-unstable parcel, pinc = 500 Pa, pseudo-liquid adiabat,
ERA5 pressure levels, 721 x 1440 columns x 37 levels.  Other workloads (same JSON contract):
C3 = calc_cape mixed-layer 500 m on HRRR-shape 1059 x 1799 x 50 model levels, C4 = calc_srh 0-3 km
on the same grid, C5 = calc_cape most-unstable on ONE time step (721 x 1440) of the 137-level
stack per rank.  Weak scaling: every rank owns one field (a time step of the stack; SURVEY §8e
shards time chunks / column blocks, no collective).

Numbers on the JSON line:
  value      whole-job columns/s with the field resident in HBM in the reference's own layout
             ([ncol, nlev] float32, level last); the timed call is the public device-pointer
             path (pres_lev_pos + relayout + kernel), CUDA events on the launch stream.
  e2e        same metric through the host-buffer C-ABI call (pinned host numpy in, host numpy
             out, H2D / D2H inside the timed region).
  roofline   the dominant kernel alone (level-major input, exactly one launch per step, CUDA
             events).  CAPE: FP-issue bound — algorithmic work = 73 flop x (moist iterations
             executed, counted by the kernel; tests assert the count equals the oracle's) against
             the FFMA peak measured on this GPU by xcape_cuda_measure_peaks, plus the HBM view.
             SRH: HBM bound — 1028 B/column against MEASURED_PEAKS.json's copy bandwidth.
  cpu_baseline  the CPU oracle (= the reference algorithm; CAPE with the libm gfortran links) on
             all host threads over a bounded sample of the same field (rank 0, N = 1).
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

UNIT = 'columns/s'
FLOP_PER_ITER = 73.0


class ClockSampler:
    """nvidia-smi sampled in the background during the timed regions (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, windows):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            inside = any(a <= ts <= b + 0.1 for a, b in windows)
            try:
                smax = float(f[2])
                if inside:
                    sm.append(float(f[1]))
            except ValueError:
                continue
            if inside:
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': smax, 'samples_under_load': len(sm),
                'reasons': sorted(reasons)}


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


class CapeWorkload:
    """calc_cape on a named synthetic configuration."""
    kind = 'cape'
    roofline_launches = 1

    precision = 'faithful'

    def __init__(self, cfg, source, ml_depth, metric, workload):
        self.cfg, self.source, self.ml_depth = cfg, source, float(ml_depth)
        self.src_id = {'surface': 1, 'most-unstable': 2, 'mixed-layer': 3}[source]
        self.metric, self.workload = metric, workload
        self.fields3 = ('t', 'td')

    def make(self, rank, cols):
        from xcape_b200.synthetic import CONFIGS, make_soundings
        kw = {}
        if self.cfg == 'C5':                   # one time step (721 x 1440) of the 24-step stack per rank
            kw['grid'] = (721, 1440)
        d = make_soundings(self.cfg, seed=CONFIGS[self.cfg]['seed'] + rank, cols=(0, cols) if cols else None,
                           winds=False, **kw)
        self.p1d = d['p'].ndim == 1
        self.roofline_launches = 2 if self.p1d else 1     # pressure grids: + the 37-thread Exner-table kernel (~2 us)
        self.ncol, self.nlev = d['t'].shape
        if not self.p1d:
            self.fields3 = ('p', 't', 'td')
        self.bytes_per_col = 4.0 * (len(self.fields3) * self.nlev) + 12 + (16 if self.src_id == 2 else 8)
        return d

    def _call(self, p, t2, td2, ps, ts, tds, plp, **kw):
        from xcape_b200.cape_cuda import cape
        kw.setdefault('precision', self.precision)
        return cape(p, t2, td2, ps, ts, tds, 1 if self.p1d else 0, plp, self.src_id, self.ml_depth, 1, 500.,
                    2 if self.p1d else 1, **kw)

    def to_device(self, d, dev):
        import torch
        return {k: torch.from_numpy(d[k]).to(dev) for k in ('p', 't', 'td', 'ps', 'ts', 'tds')}

    def step_dev(self, g):
        p = g['p'] if self.p1d else g['p'].t()
        return self._call(p, g['t'].t(), g['td'].t(), g['ps'], g['ts'], g['tds'], None)

    def kernel_state(self, g):
        from xcape_b200.cape_cuda import pres_lev_pos
        st = dict(t=g['t'].t().contiguous(), td=g['td'].t().contiguous(),
                  p=g['p'] if self.p1d else g['p'].t().contiguous(),
                  plp=pres_lev_pos(g['p'], g['ps']) if self.p1d else 1)
        # algorithmic work = iterations of the REFERENCE algorithm: the faithful kernel's count,
        # which tests assert equal to the oracle's; the fast modes execute fewer passes for the same answer
        cnt = self._call(st['p'], st['t'], st['td'], g['ps'], g['ts'], g['tds'], st['plp'], return_counters=True,
                         precision='faithful')
        st['total_iter'] = float(cnt[5].double().sum().item())
        if self.precision != 'faithful':
            cnt = self._call(st['p'], st['t'], st['td'], g['ps'], g['ts'], g['tds'], st['plp'], return_counters=True)
        st['executed_iter'] = float(cnt[5].double().sum().item())
        return st

    def step_kernel(self, g, st, **kw):
        return self._call(st['p'], st['t'], st['td'], g['ps'], g['ts'], g['tds'], st['plp'], **kw)

    def pinned(self, d):
        import torch
        keys = self.fields3 + ('ps', 'ts', 'tds')
        pin = {k: torch.empty(d[k].shape, dtype=torch.float32).pin_memory() for k in keys}
        for k in keys:
            pin[k].numpy()[...] = d[k]
        hp = {k: v.numpy() for k, v in pin.items()}
        hp['_keep'] = pin
        if self.p1d:
            hp['p'] = d['p']
        return hp

    def step_e2e(self, hp, device):
        p = hp['p'] if self.p1d else hp['p'].T
        return self._call(p, hp['t'].T, hp['td'].T, hp['ps'], hp['ts'], hp['tds'], None, device=device)

    def io_bytes(self):
        h2d = int(len(self.fields3) * self.ncol * self.nlev * 4 + 3 * self.ncol * 4 + (self.nlev * 4 if self.p1d else 0))
        return h2d, int(16 * self.ncol)

    def cpu_rate(self, d, nsample, nthreads, counters=False):
        import oracle
        idx = np.arange(0, self.ncol, max(1, self.ncol // nsample))[:nsample]
        sub = {k: np.ascontiguousarray(d[k][idx]) for k in ('t', 'td', 'ps', 'ts', 'tds')}
        p = d['p'] if self.p1d else np.ascontiguousarray(d['p'][idx])
        t0 = time.perf_counter()
        out = oracle.calc_cape_ref(p, sub['t'], sub['td'], sub['ps'], sub['ts'], sub['tds'], source=self.source,
                                   ml_depth=self.ml_depth, pinc=500., adiabat='pseudo-liquid',
                                   vertical_lev='pressure' if self.p1d else 'sigma', tmode=oracle.LIBM,
                                   nthreads=nthreads, counters=counters)
        dt = time.perf_counter() - t0
        note = f"; {float(out[1]['n_iter'].mean()):.1f} iterations/column" if counters else ''
        return len(idx) / dt, len(idx), 'oracle tmode=LIBM (reference algorithm, glibc libm)' + note

    def roofline(self, ms_kernel, st, fp32_peak, fp64_peak, hbm_peak, hbm_src):
        tf = FLOP_PER_ITER * st['total_iter'] / (ms_kernel * 1e-3) / 1e12
        gbs = self.bytes_per_col * self.ncol / (ms_kernel * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic(f'cape_{self.cfg}_{self.precision}', self.ncol)
        return {
            'bound': 'fp32', 'achieved': tf, 'peak': fp32_peak, 'unit': 'TFLOP/s', 'frac': tf / fp32_peak,
            'traffic': traffic, 'traffic_unit': 'DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, scaled by columns)',
            'traffic_source': traffic_src,
            'kernel': f'cape_kernel<MathSpec,{self.src_id},1,{str(self.p1d).lower()}>', 'kernel_ms': ms_kernel,
            'work': f"{FLOP_PER_ITER:.0f} flop x {st['total_iter'] / self.ncol:.1f} moist iterations/column of the reference algorithm "
                    f"(= the faithful kernel's count); this kernel executed {st['executed_iter'] / self.ncol:.1f}/column",
            'peak_source': 'FFMA microbenchmark on this GPU (xcape_cuda_measure_peaks), 2 flop/FMA',
            'fp64_peak_tflops': fp64_peak, 'iterations_per_s': st['total_iter'] / (ms_kernel * 1e-3),
            'hbm': {'bound': 'hbm', 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbs / hbm_peak,
                    'bytes_per_column': self.bytes_per_col, 'peak_source': hbm_src}}


class SrhWorkload:
    """calc_srh (0-3 km, Bunkers storm motion) on a named synthetic configuration."""
    kind = 'srh'
    roofline_launches = 2        # streaming kernel + the (normally empty, ~3 us) EXACT work-list kernel
    KEYS3 = ('p', 't', 'td', 'u', 'v')
    KEYS1 = ('ps', 'ts', 'tds', 'us', 'vs')
    precision = 'faithful'

    def __init__(self, cfg, metric, workload):
        self.cfg, self.metric, self.workload = cfg, metric, workload

    def make(self, rank, cols):
        from xcape_b200.synthetic import CONFIGS, make_soundings
        d = make_soundings(self.cfg, seed=CONFIGS[self.cfg]['seed'] + rank, cols=(0, cols) if cols else None)
        self.ncol, self.nlev = d['t'].shape
        self.bytes_per_col = 4.0 * (5 * self.nlev) + 20 + 8
        return d

    def _call(self, a, output=1, **kw):
        from xcape_b200.srh_cuda import srh_fused
        kw.setdefault('precision', self.precision)
        return srh_fused(a['p'], a['t'], a['td'], a['u'], a['v'], a['ps'], a['ts'], a['tds'], a['us'], a['vs'],
                         0, None, 3000, 2., 1, output, **kw)

    def to_device(self, d, dev):
        import torch
        return {k: torch.from_numpy(d[k]).to(dev) for k in self.KEYS3 + self.KEYS1}

    def step_dev(self, g):
        a = {k: g[k].t() for k in self.KEYS3}
        a.update({k: g[k] for k in self.KEYS1})
        return self._call(a)

    def kernel_state(self, g):
        st = {k: g[k].t().contiguous() for k in self.KEYS3}
        st.update({k: g[k] for k in self.KEYS1})
        return st

    def step_kernel(self, g, st):
        return self._call(st)

    def pinned(self, d):
        import torch
        pin = {k: torch.empty(d[k].shape, dtype=torch.float32).pin_memory() for k in self.KEYS3 + self.KEYS1}
        for k in pin:
            pin[k].numpy()[...] = d[k]
        hp = {k: v.numpy() for k, v in pin.items()}
        hp['_keep'] = pin
        return hp

    def step_e2e(self, hp, device):
        a = {k: hp[k].T for k in self.KEYS3}
        a.update({k: hp[k] for k in self.KEYS1})
        return self._call(a, device=device)

    def io_bytes(self):
        return int(5 * self.ncol * self.nlev * 4 + 5 * self.ncol * 4), int(16 * self.ncol)

    def cpu_rate(self, d, nsample, nthreads, counters=False):
        import oracle
        idx = np.arange(0, self.ncol, max(1, self.ncol // nsample))[:nsample]
        sub = [np.ascontiguousarray(d[k][idx]) for k in self.KEYS3 + self.KEYS1]
        t0 = time.perf_counter()
        oracle.calc_srh_ref(*sub, depth=3000, vertical_lev='sigma', output_var='srh', nthreads=nthreads)
        dt = time.perf_counter() - t0
        return len(idx) / dt, len(idx), 'oracle stdheight+Bunkers+SREH chain (reference algorithm)'

    def roofline(self, ms_kernel, st, fp32_peak, fp64_peak, hbm_peak, hbm_src):
        gbs = self.bytes_per_col * self.ncol / (ms_kernel * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic('srh_' + self.cfg, self.ncol)
        return {'bound': 'hbm', 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbs / hbm_peak,
                'traffic': traffic, 'traffic_source': traffic_src,
                'kernel': 'srh_kernel<float,false,false> (+ srh_exact_kernel on an empty work list)', 'kernel_ms': ms_kernel, 'bytes_per_column': self.bytes_per_col,
                'peak_source': hbm_src, 'fp64_peak_tflops': fp64_peak,
                'note': 'faithful: hypsometric exp/log chain in binary64 (reference arithmetic); fast: binary32'}


def ncu_traffic(key, ncol):
    """DRAM bytes per launch from the committed ncu capture of the same kernel (profiles/ncu_traffic.json)."""
    try:
        t = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))[key]
        return t['dram_bytes_per_column'] * ncol, t['source']
    except (OSError, KeyError, ValueError):
        return None, None


def get_workload(name):
    if name == 'C2':
        return CapeWorkload('C2', 'most-unstable', 500., 'CAPE/CIN columns/sec (ERA5-shape, MU parcel)',
                            'calc_cape most-unstable, ERA5 pressure levels 721x1440x37, pinc=500 Pa, pseudo-liquid (configs[1])')
    if name == 'C3':
        return CapeWorkload('C3', 'mixed-layer', 500., 'CAPE/CIN columns/sec (HRRR-shape, ML parcel)',
                            'calc_cape mixed-layer (ml_depth=500 m), HRRR model levels 1059x1799x50, pinc=500 Pa (configs[2])')
    if name == 'C4':
        return SrhWorkload('C4', 'SRH columns/sec (HRRR-shape, 0-3 km, Bunkers)',
                           'calc_srh 0-3 km with Bunkers storm motion, HRRR model levels 1059x1799x50 (configs[3])')
    if name == 'C5':
        return CapeWorkload('C5', 'most-unstable', 500., 'CAPE/CIN columns/sec (ERA5 137 model levels, MU parcel)',
                            'calc_cape most-unstable, one 721x1440x137 time step of the 24-step stack per GPU (configs[4])')
    raise SystemExit(f'unknown workload {name}')


def run_reference(args, rank, world, wl):
    """--impl reference: the reference's CPU algorithm on this box's host cores."""
    if rank != 0:
        return
    import oracle
    oracle.build()
    nth = host_threads()
    d = wl.make(0, args.cols)
    rate0, _, _ = wl.cpu_rate(d, min(wl.ncol, 4000 * nth), nth)
    sample = int(min(wl.ncol, max(2000 * nth, rate0 * args.ref_seconds)))
    for _ in range(args.warmup):
        wl.cpu_rate(d, sample, nth)
    t0 = time.perf_counter()
    n_done = 0
    for _ in range(args.steps):
        _, n, how = wl.cpu_rate(d, sample, nth)
        n_done += n
    dt = time.perf_counter() - t0
    v = n_done / dt
    desc = f'{n} of {wl.ncol} columns per step (evenly strided), {how}, {nth} threads'
    print(json.dumps({
        'impl': 'reference', 'metric': wl.metric, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': wl.workload, 'sample': desc},
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': nth, 'kind': 'port', 'sample': desc},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='cuda', choices=['cuda', 'reference'])
    ap.add_argument('--workload', default='C2', choices=['C2', 'C3', 'C4', 'C5'])
    ap.add_argument('--cols', type=int, default=0, help='debug: use only the first COLS columns of the field')
    ap.add_argument('--cpu-seconds', type=float, default=12.0, help='CPU work budget of the cpu_baseline leg')
    ap.add_argument('--ref-seconds', type=float, default=3.0, help='CPU seconds per step of --impl reference')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--precision', default='faithful', choices=['faithful', 'fast'],
                    help="CAPE arithmetic of every timed leg (default: faithful = bit-exact vs the oracle's SPEC mode)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'cuda' else args.warmup

    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    wl = get_workload(args.workload)
    wl.precision = args.precision

    if args.impl == 'reference':
        return run_reference(args, rank, world, wl)

    # stdout must carry exactly ONE JSON line: NCCL prints "NCCL version ..." to the C-level stdout when the
    # first communicator is created, so fd 1 points at stderr until the line is ready.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from xcape_b200 import _lib

    if not torch.cuda.is_available() or _lib.device_count() < 1:
        raise SystemExit('bench.py: no CUDA device — the CUDA path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    numa_node = None
    full_affinity = os.sched_getaffinity(0) if hasattr(os, 'sched_getaffinity') else None
    if not os.environ.get('XCAPE_BENCH_NO_NUMA_BIND'):
        from xcape_b200.sharding import bind_host_to_gpu
        numa_node = bind_host_to_gpu(local_rank)      # before any pinned allocation of this rank
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- this rank's field: time step `rank` of the synthetic stack -------------------------
    d = wl.make(rank, args.cols)
    ncol, nlev = wl.ncol, wl.nlev
    g = wl.to_device(d, dev)       # device-resident, reference layout ([ncol, nlev], level last)
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get('XCAPE_BENCH_NO_SAMPLER'):
        # nvidia-smi's start-up (NVML init) stalls the GPU for tens of ms: let it reach its steady
        # sampling state before anything is timed
        sampler.start()
        t_wait = time.time()
        while not sampler.rows and time.time() - t_wait < 5.0:
            time.sleep(0.05)
    windows = []

    # Warm up with the same object lifetimes as the timed loop (`out_dev = ...` keeps one step's
    # outputs alive while the next step allocates): otherwise torch's caching allocator meets that
    # pattern for the first time inside the timed region and its cudaMalloc blocks the enqueuing
    # thread for 20-80 ms in the second timed step (profiles/probe_step1_stall.py).
    out_dev = None
    for _ in range(args.warmup):
        out_dev = wl.step_dev(g)
    barrier()
    t_warm = time.time()
    while time.time() - t_warm < 0.5:      # clocks / power state settled (untimed)
        out_dev = wl.step_dev(g)
    barrier()
    # Python's cyclic GC can pause the enqueuing thread for 40-90 ms in the middle of a step (seen as
    # one slow step in ~1 of 3 runs); like timeit, collect first and keep it off while timing.
    gc.collect()
    gc.disable()
    l0 = _lib.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e0.record()
    for i in range(args.steps):
        out_dev = wl.step_dev(g)
        marks[i].record()
    e1.record()
    barrier()
    windows.append((w0, time.time()))
    launches = _lib.kernel_launches() - l0
    per_step = [a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks)]
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    value = world * ncol / (ms_step * 1e-3)

    # ---- roofline leg: the dominant kernel alone (level-major input, one launch per step) -----
    st = wl.kernel_state(g)
    for _ in range(3):
        wl.step_kernel(g, st)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lk = _lib.kernel_launches()
    w0 = time.time()
    k0.record()
    for _ in range(args.steps):
        wl.step_kernel(g, st)
    k1.record()
    torch.cuda.synchronize()
    windows.append((w0, time.time()))
    assert _lib.kernel_launches() - lk == args.steps * wl.roofline_launches, 'roofline leg: unexpected kernel count'
    ms_kernel = k0.elapsed_time(k1) / args.steps
    other = None
    if wl.kind == 'cape':        # the other precision mode, for the record
        alt = 'fast' if args.precision == 'faithful' else 'faithful'
        for _ in range(3):
            wl.step_kernel(g, st, precision=alt)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(args.steps):
            out_alt = wl.step_kernel(g, st, precision=alt)
        a1.record()
        torch.cuda.synchronize()
        ms_alt = a0.elapsed_time(a1) / args.steps
        ref_k = wl.step_kernel(g, st)
        dc = (out_alt[0] - ref_k[0]).abs()
        lim = torch.clamp(1e-4 * ref_k[0].abs(), min=1.0)
        lim_i = torch.clamp(1e-4 * ref_k[1].abs(), min=1.0)
        wl.precision = alt
        for _ in range(3):
            wl.step_dev(g)
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(args.steps):
            wl.step_dev(g)
        a1.record()
        torch.cuda.synchronize()
        ms_alt_dev = a0.elapsed_time(a1) / args.steps
        wl.precision = args.precision
        other = {'precision': alt, 'kernel_ms': ms_alt, 'kernel_columns_per_s': ncol / (ms_alt * 1e-3),
                 'device_path_ms_per_step': ms_alt_dev, 'device_path_columns_per_s_this_rank': ncol / (ms_alt_dev * 1e-3),
                 'vs_' + args.precision: {'columns': ncol,
                                          'cape_or_cin_outside_max(1,1e-4rel)': int(((dc > lim) | ((out_alt[1] - ref_k[1]).abs() > lim_i)).sum().item()),
                                          'max_abs_dcape': float(dc.max().item()), 'mean_abs_dcape': float(dc.mean().item()),
                                          'mulev_differs': int((out_alt[2] != ref_k[2]).sum().item())}}
        del out_alt, ref_k
    st = {k: v for k, v in st.items() if not hasattr(v, 'is_cuda')}     # drop the level-major copies

    # ---- end to end: pinned host buffers through the host-pointer C-ABI call ------------------
    hp = wl.pinned(d)
    for _ in range(2):
        out_host = wl.step_e2e(hp, local_rank)
    barrier()
    w0 = time.time()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_host = wl.step_e2e(hp, local_rank)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    barrier()
    windows.append((w0, time.time()))
    e2e_value = world * ncol / e2e_s
    # Two host threads issuing alternate steps (what dask's threaded scheduler does with GIL-releasing
    # calls): the second call's H2D overlaps the first call's tail.  Reported next to the one-thread e2e.
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(2) as ex:
        list(ex.map(lambda _: wl.step_e2e(hp, local_rank), range(2)))
        barrier()
        t0 = time.perf_counter()
        list(ex.map(lambda _: wl.step_e2e(hp, local_rank), range(args.steps)))
        e2e2_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    barrier()
    if other is not None:
        wl.precision = other['precision']
        wl.step_e2e(hp, local_rank)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            wl.step_e2e(hp, local_rank)
        other['e2e_ms_per_step'] = (time.perf_counter() - t0) / args.steps * 1e3
        other['e2e_columns_per_s_this_rank'] = ncol / (other['e2e_ms_per_step'] * 1e-3)
        wl.precision = args.precision
    h2d, d2h = wl.io_bytes()
    same = all(np.array_equal(np.asarray(a), b.cpu().numpy()) for a, b in zip(out_host, out_dev))

    gc.enable()
    clocks = sampler.stop(windows) if (rank == 0 and sampler.proc is not None) else None

    # ---- peaks + CPU baseline (rank 0) ---------------------------------------------------------
    if rank == 0:
        fp32_peak, fp64_peak = _lib.measure_peaks(local_rank, 5)
        peaks_file = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        hbm_peak, hbm_src = 6650.0, 'fallback (B200_PROFILING.md)'
        if os.path.exists(peaks_file):
            hbm_peak, hbm_src = float(json.load(open(peaks_file))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        roofline = wl.roofline(ms_kernel, st, fp32_peak, fp64_peak, hbm_peak, hbm_src)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            import oracle
            oracle.build()
            if full_affinity is not None:
                os.sched_setaffinity(0, full_affinity)      # the CPU baseline gets every host core again
            nth = host_threads()
            r0, _, _ = wl.cpu_rate(d, min(ncol, 2000 * nth), nth)
            sample = int(min(ncol, max(1000 * nth, r0 * args.cpu_seconds)))
            rate, n, how = wl.cpu_rate(d, sample, nth, counters=True)
            cpu = {'value': rate, 'unit': UNIT, 'cores': nth, 'kind': 'port',
                   'sample': f'{n} of {ncol} columns (evenly strided) of the same field, {how}, {nth} threads'}
        line = {
            'metric': wl.metric, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {'workload': wl.workload, 'columns_per_gpu': ncol, 'levels': nlev, 'layout': 'level-last [ncol, nlev] float32 '
                       '(reference layout), resident in HBM', 'precision': args.precision + (' (CAPE bit-exact vs oracle SPEC arithmetic)' if args.precision == 'faithful' else
                                                     ' (FP32-pipe moist body; tolerance-level parity, MU level exact)'),
                       'l2': f'inputs {h2d / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)',
                       'parallelism': f'{world} x independent column shards, no collective',
                       'host_numa_node_of_rank0': numa_node},
            'ms_per_step_minmedmax': [float(np.min(per_step)), float(np.median(per_step)), float(np.max(per_step))],
            'ms_each_step': [round(float(x), 3) for x in per_step],
            'clocks': clocks, 'gpu_launches': int(launches),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': e2e_s * 1e3, 'matches_device_path': bool(same),
                    'two_host_threads': {'value': world * ncol / e2e2_s, 'ms_per_step': e2e2_s * 1e3}},
            'roofline': roofline, 'cpu_baseline': cpu, 'other_precision': other}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
