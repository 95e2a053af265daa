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
             path (pres_lev_pos + the sorted call reading the layout in place: source-parcel kernel, ordering, ascent kernel), CUDA events on the launch stream.
  e2e        same metric through the PUBLIC call a user makes, core.calc_cape / core.calc_srh(method='cuda'),
             on pinned host numpy arrays in the reference's layout ([ncol, nlev], level last): host numpy in,
             host numpy out, H2D / D2H inside the timed region.  Sub-records: `pageable` (ordinary numpy
             arrays), `level_major_pinned` ([nlev, ncol] = the on-disk order of ERA5 / HRRR files, lev_axis=0:
             zero relayout, and on pressure grids only the levels the ascent can reach are copied).
  roofline   the dominant kernel alone (level-major input, exactly one launch per step, CUDA
             events).  CAPE: FP-issue bound — algorithmic work = 73 flop x (moist iterations
             executed, counted by the kernel; tests assert the count equals the oracle's) against
             the FFMA peak measured on this GPU by xcape_cuda_measure_peaks, plus the HBM view.
             SRH: HBM bound — 1028 B/column against MEASURED_PEAKS.json's copy bandwidth.
  cpu_baseline  the CPU oracle (= the reference algorithm; CAPE with the libm gfortran links) on
             all host threads AND on one thread over bounded samples of the same field, 3 repeats each,
             best and median (rank 0, N = 1).
  kernel_variants  the same kernel on the column-shuffled field (warp-divergence worst case) and on the
             global-mix field (~45 % of columns gated by ts <= 0 degC).
  workloads  (N = 1) C3 / C4 / C5 sub-records: kernel ms, device-path ms, columns/s, roofline fraction.
  strong_scaling_C5  BASELINE configs[4]: the 24-step 721x1440x137 stack, 24/N steps per rank, whole-job
             columns/s device-resident and end to end (each rank re-uses one synthetic step's buffers).
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
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12      # 148 SMs x 128 FP32 lanes x 2 flop/FMA x 1.965 GHz = 74.4


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

    def step_e2e(self, hp, device, lev_axis=-1):
        """The public API on host arrays ([ncol, nlev], or [nlev, ncol] with lev_axis=0)."""
        from xcape_b200 import core
        return core.calc_cape(hp['p'], hp['t'], hp['td'], hp['ps'], hp['ts'], hp['tds'], source=self.source,
                              ml_depth=self.ml_depth, adiabat='pseudo-liquid', pinc=500., method='cuda',
                              vertical_lev='pressure' if self.p1d else 'sigma', device=device, precision=self.precision,
                              lev_axis=lev_axis)

    def level_major(self, hp, pin=True):
        """[nlev, ncol] copies of the 3-D host fields (pinned): the on-disk order of ERA5 / HRRR."""
        import torch
        out = dict(hp)
        keep = []
        for k in self.fields3:
            a = torch.empty((self.nlev, self.ncol), dtype=torch.float32)
            a = a.pin_memory() if pin else a
            a.numpy()[...] = hp[k].T
            keep.append(a)
            out[k] = a.numpy()
        out['_keep_lm'] = keep
        return out

    def shipped_levels(self):
        """Levels the host path copies for level-major / pageable input on a pressure grid (api.cu)."""
        if not self.p1d:
            return self.nlev
        from xcape_b200.synthetic import ERA5_LEVELS_HPA
        below = np.flatnonzero(ERA5_LEVELS_HPA[:self.nlev] <= 100.0)
        need = int(below[0]) + 1 if below.size else self.nlev
        return need if need + 2 <= self.nlev else self.nlev

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
            'peak_nominal': NOMINAL_FP32_TFLOPS, 'frac_of_nominal': tf / NOMINAL_FP32_TFLOPS,
            'traffic': traffic, 'traffic_unit': 'DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, scaled by columns)',
            'traffic_source': traffic_src,
            'kernel': (f'cape_kernel2<MathSpec,1,1,{str(self.p1d).lower()},true> (two columns per thread, packed FP32, sorted execution)'
                       if self.precision == 'faithful' else f'cape_kernel<MathFast,{self.src_id},1,{str(self.p1d).lower()}>'), 'kernel_ms': ms_kernel,
            'work': f"{FLOP_PER_ITER:.0f} flop x {st['total_iter'] / self.ncol:.1f} moist iterations/column of the reference algorithm "
                    f"(= the faithful kernel's count); this kernel executed {st['executed_iter'] / self.ncol:.1f}/column",
            'peak_source': 'FFMA microbenchmark on this GPU (xcape_cuda_measure_peaks), 2 flop/FMA',
            'call_ms': st.get('call_ms'), 'launches_per_call': st.get('launches_per_call'),
            'frac_of_call': (FLOP_PER_ITER * st['total_iter'] / (st['call_ms'] * 1e-3) / 1e12 / fp32_peak) if st.get('call_ms') else None,
            'call': 'the device call on level-major input: source-parcel kernel + column ordering + ascent kernel (faithful); kernel_ms / achieved / frac are the ascent kernel alone, timed by an event pair the library records around it',
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

    def step_e2e(self, hp, device, lev_axis=-1):
        from xcape_b200 import core
        return core.calc_srh(*(hp[k] for k in self.KEYS3 + self.KEYS1), depth=3000, vertical_lev='sigma', output_var='srh',
                             method='cuda', device=device, precision=self.precision, lev_axis=lev_axis)

    fields3 = KEYS3

    def level_major(self, hp, pin=True):
        return CapeWorkload.level_major(self, hp, pin)

    def shipped_levels(self):
        return self.nlev

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
                'kernel': 'srh_kernel<float,false,false,false> (level-major input; + srh_exact_kernel on an empty work list in call_ms)', 'kernel_ms': ms_kernel, 'bytes_per_column': self.bytes_per_col,
                'call_ms': st.get('call_ms'), 'launches_per_call': st.get('launches_per_call'),
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
    ap.add_argument('--cpu-seconds', type=float, default=18.0, help='CPU work budget of the cpu_baseline leg (6 repeats in all)')
    ap.add_argument('--ref-seconds', type=float, default=3.0, help='CPU seconds per step of --impl reference')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip kernel variants, the C3/C4/C5 sub-records and the C5 stack')
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

    def time_device(fn, warm, steps):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(steps):
            fn()
        a1.record()
        torch.cuda.synchronize()
        return a0.elapsed_time(a1) / steps

    def time_kernel_alone(fn, steps):
        """Mean device time of the dominant kernel of fn()'s call (event pair recorded by the library around it)."""
        _lib.time_kernels(True)
        ks = []
        for _ in range(steps):
            fn()
            ks.append(_lib.last_kernel_ms())
        _lib.time_kernels(False)
        return float(np.mean(ks))
    time_device.kernel_alone = time_kernel_alone

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
    launches_per_call = (_lib.kernel_launches() - lk) / args.steps
    ms_call = k0.elapsed_time(k1) / args.steps       # every kernel of the call on level-major input (source parcels, ordering, ascent)
    # the dominant kernel alone: the library records a CUDA event pair around it on the launch stream
    _lib.time_kernels(True)
    w0 = time.time()
    ks = []
    for _ in range(args.steps):
        wl.step_kernel(g, st)
        ks.append(_lib.last_kernel_ms())
    _lib.time_kernels(False)
    windows.append((w0, time.time()))
    ms_kernel = float(np.mean(ks))
    st['call_ms'] = ms_call
    st['launches_per_call'] = launches_per_call
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

    # ---- kernel on the adversarial / mixed fields (CAPE only): same kernel, same launch shape ---------
    variants = None
    if wl.kind == 'cape' and not args.no_extras and not args.cols:
        variants = {}
        for name, kw in (('shuffled_active', dict(shuffle=True)), ('global_mix', dict(active=False))):
            from xcape_b200.synthetic import CONFIGS, make_soundings
            dv = make_soundings(wl.cfg, seed=CONFIGS[wl.cfg]['seed'] + rank, winds=False,
                                **({'grid': (721, 1440)} if wl.cfg == 'C5' else {}), **kw)
            gv = wl.to_device(dv, dev)
            stv = wl.kernel_state(gv)
            ms_vc = time_device(lambda: wl.step_kernel(gv, stv), 3, args.steps)
            ms_v = time_kernel_alone(lambda: wl.step_kernel(gv, stv), args.steps)
            variants[name] = {'kernel_ms': ms_v, 'call_ms': ms_vc, 'columns_per_s': ncol / (ms_vc * 1e-3),
                              'reference_iterations_per_column': stv['total_iter'] / ncol,
                              'frac_of_fp32_peak_measured_below': None}
            if name == 'global_mix':
                variants[name]['gated_fraction'] = float((dv['ts'] <= 0).mean())
            del gv, stv, dv
        torch.cuda.empty_cache()

    # ---- end to end: the public API (core.calc_cape / calc_srh, method='cuda') on host arrays -----------
    hp = wl.pinned(d)

    def time_host(fn, warm, steps):
        for _ in range(warm):
            out = fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = fn()
        torch.cuda.synchronize()
        dt = max_over_ranks((time.perf_counter() - t0) / steps)
        barrier()
        return dt, out

    w0 = time.time()
    e2e_s, out_host = time_host(lambda: wl.step_e2e(hp, local_rank), max(3, args.warmup), args.steps)
    windows.append((w0, time.time()))
    e2e_value = world * ncol / e2e_s
    h2d, d2h = wl.io_bytes()
    same = all(np.array_equal(np.asarray(a), b.cpu().numpy()) for a, b in zip(out_host, out_dev))
    e2e_extra = {}
    # ordinary (pageable) numpy arrays: what a caller who knows nothing about pinned memory passes
    dpage = {k: np.array(d[k], copy=True) for k in wl.fields3 + ('ps', 'ts', 'tds') + (('us', 'vs') if wl.kind == 'srh' else ())}
    if wl.kind == 'cape' and wl.p1d:
        dpage['p'] = d['p']
    pg_s, out_pg = time_host(lambda: wl.step_e2e(dpage, local_rank), 3, max(3, args.steps // 2))
    e2e_extra['pageable'] = {'value': world * ncol / pg_s, 'ms_per_step': pg_s * 1e3,
                             'matches': bool(all(np.array_equal(a, b) for a, b in zip(out_pg, out_host)))}
    del dpage
    # level-major pinned arrays: the on-disk order (time, level, lat, lon) of ERA5 / HRRR files
    hlm = wl.level_major(hp)
    lm_s, out_lm = time_host(lambda: wl.step_e2e(hlm, local_rank, lev_axis=0), 3, args.steps)
    ship = wl.shipped_levels()
    e2e_extra['level_major_pinned'] = {
        'value': world * ncol / lm_s, 'ms_per_step': lm_s * 1e3, 'levels_copied': ship, 'levels': nlev,
        'h2d_bytes_per_step': int(h2d - (nlev - ship) * ncol * 4 * (2 if (wl.kind == 'cape' and wl.p1d) else 0)),
        'matches': bool(all(np.array_equal(a, b) for a, b in zip(out_lm, out_host)))}
    del hlm
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
        o_s, _ = time_host(lambda: wl.step_e2e(hp, local_rank), 3, args.steps)
        other['e2e_ms_per_step'] = o_s * 1e3
        other['e2e_columns_per_s'] = world * ncol / o_s
        wl.precision = args.precision
    del hp

    # ---- BASELINE configs[4]: the 24-step C5 stack, 24/N steps per rank (strong scaling) ---------------
    strong = None
    if not args.no_extras and not args.cols and 24 % world == 0:
        strong = strong_scaling_c5(rank, local_rank, world, dev, barrier, max_over_ranks, time_device, args)

    # ---- the other named workloads as sub-records (N = 1) --------------------------------------------------
    extras = None
    if world == 1 and not args.no_extras and not args.cols and args.workload == 'C2':
        extras = {}
        shared = {}
        for name in ('C3', 'C4', 'C5'):
            extras[name] = sub_record(name, dev, local_rank, time_device, shared, args)
            torch.cuda.empty_cache()

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
        if wl.kind == 'cape':
            rrr = _lib.measure_fp32_rrr(local_rank, 5)
            roofline['fp32_peak_three_register_operands'] = {
                'peak': rrr, 'frac': roofline['achieved'] / rrr,
                'what': 'FFMA with three distinct register operands and no reuse (xcape_cuda_measure_fp32_rrr): the rate the '
                        'register file sustains for real code; `peak` above feeds two operands from uniform registers'}
        if other is not None and 'total_iter' in st:
            tf_o = FLOP_PER_ITER * st['total_iter'] / (other['kernel_ms'] * 1e-3) / 1e12
            other['roofline'] = {'bound': 'fp32', 'achieved': tf_o, 'peak': fp32_peak, 'unit': 'TFLOP/s', 'frac': tf_o / fp32_peak,
                                 'work': 'reference-algorithm iterations (the fast solve executes about a third of them)'}
            other['value'] = world * ncol / (other['device_path_ms_per_step'] * 1e-3)
            other['parity'] = ('tolerance-level: every well-conditioned column inside max(1 J/kg, 1e-4 rel) for CAPE and CIN, MU level and '
                               'convergence status identical (tests/test_gpu_parity.py::test_fast_mode_full_era5_field)')
        if variants:
            for v in variants.values():
                v['frac_of_fp32_peak'] = FLOP_PER_ITER * v['reference_iterations_per_column'] * ncol / (v['kernel_ms'] * 1e-3) / 1e12 / fp32_peak
                v.pop('frac_of_fp32_peak_measured_below')
        if extras:
            for name, r in extras.items():
                if 'reference_iterations_per_column' in r:
                    tf = FLOP_PER_ITER * r['reference_iterations_per_column'] * r['columns'] / (r['kernel_ms'] * 1e-3) / 1e12
                    r['roofline'] = {'bound': 'fp32', 'achieved': tf, 'peak': fp32_peak, 'unit': 'TFLOP/s', 'frac': tf / fp32_peak}
                else:
                    gbs = r['bytes_per_column'] * r['columns'] / (r['kernel_ms'] * 1e-3) / 1e9
                    gbs_dev = r['bytes_per_column'] * r['columns'] / (r['device_path_ms_per_step'] * 1e-3) / 1e9
                    r['roofline'] = {'bound': 'hbm', 'achieved': gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': gbs / hbm_peak,
                                     'device_path_achieved': gbs_dev, 'device_path_frac': gbs_dev / hbm_peak}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline(wl, d, full_affinity, args)
        line = {
            'metric': wl.metric, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {'workload': wl.workload, 'columns_per_gpu': ncol, 'levels': nlev, 'layout': 'level-last [ncol, nlev] float32 '
                       '(reference layout), resident in HBM', 'precision': args.precision + (' (CAPE bit-exact vs oracle SPEC arithmetic)' if args.precision == 'faithful' else
                                                     ' (FP32-pipe moist body; tolerance-level parity, MU level exact)'),
                       'l2': f'inputs {h2d / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)',
                       'parallelism': f'{world} x independent column shards, no collective',
                       'host_cpu_binding_of_rank0': numa_node},
            'ms_per_step_minmedmax': [float(np.min(per_step)), float(np.median(per_step)), float(np.max(per_step))],
            'ms_each_step': [round(float(x), 3) for x in per_step],
            'clocks': clocks, 'gpu_launches': int(launches),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': e2e_s * 1e3, 'matches_device_path': bool(same),
                    'api': "core.calc_cape(method='cuda')" if wl.kind == 'cape' else "core.calc_srh(method='cuda')",
                    'input': 'pinned host numpy, reference layout [ncol, nlev]',
                    'h2d_gb_per_s_per_rank': h2d / e2e_s / 1e9,
                    'two_host_threads': {'value': world * ncol / e2e2_s, 'ms_per_step': e2e2_s * 1e3}, **e2e_extra},
            'roofline': roofline, 'cpu_baseline': cpu, 'other_precision': other, 'kernel_variants': variants,
            'workloads': extras, 'strong_scaling_C5': strong}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(wl, d, full_affinity, args):
    """The reference algorithm (CPU oracle, glibc libm) on all host threads and on one thread: 3 repeats each on a
    bounded, evenly strided sample of the same field; best and median."""
    import oracle
    oracle.build()
    if full_affinity is not None:
        os.sched_setaffinity(0, full_affinity)      # the CPU baseline gets every host core again
    nth = host_threads()
    r0, _, _ = wl.cpu_rate(d, min(wl.ncol, 2000 * nth), nth)
    per_rep = args.cpu_seconds / 6.0                  # 3 + 3 repeats
    out = {}
    for label, threads, rate_guess in (('all_threads', nth, r0), ('one_thread', 1, r0 / nth)):
        sample = int(min(wl.ncol, max(500 * threads, rate_guess * per_rep)))
        rates, how, n = [], '', 0
        for rep in range(3):
            rate, n, how = wl.cpu_rate(d, sample, threads, counters=(rep == 0))
            rates.append(rate)
        out[label] = {'best': float(np.max(rates)), 'median': float(np.median(rates)), 'threads': threads,
                      'columns_per_repeat': n, 'repeats': 3, 'how': how}
    best = out['all_threads']
    return {'value': best['best'], 'unit': UNIT, 'cores': nth, 'kind': 'port',
            'sample': f"{best['columns_per_repeat']} of {wl.ncol} columns (evenly strided) of the same field, best of 3 repeats, "
                      f"{best['how']}, {nth} threads",
            'median': best['median'], 'one_thread': out['one_thread'],
            'why_port': 'the reference is Fortran: no Fortran compiler here or on the GPU box (probed: gfortran / flang / nvfortran / ifx '
                        'absent, only the libgfortran runtime inside numpy/scipy wheels), so the CPU arm is the line-by-line C++ oracle'}


def sub_record(name, dev, local_rank, time_device, shared, args):
    """One of the other BASELINE workloads on this GPU: kernel alone, device path in the reference layout, e2e."""
    import torch
    wl = get_workload(name)
    wl.precision = 'faithful'
    from xcape_b200.synthetic import CONFIGS, make_soundings
    if name in ('C3', 'C4'):                       # same grid and seed: generate once, with winds
        if 'hrrr' not in shared:
            shared['hrrr'] = make_soundings('C4', seed=CONFIGS['C4']['seed'])
        d = shared['hrrr']
        wl.ncol, wl.nlev = d['t'].shape
        if wl.kind == 'cape':
            wl.p1d = False
            wl.fields3 = ('p', 't', 'td')
            wl.roofline_launches = 1
            wl.bytes_per_col = 4.0 * (3 * wl.nlev) + 12 + 8
        else:
            wl.bytes_per_col = 4.0 * (5 * wl.nlev) + 20 + 8
    else:
        d = wl.make(0, 0)
    g = wl.to_device(d, dev)
    steps = max(3, min(args.steps, 5))
    ms_dev = time_device(lambda: wl.step_dev(g), 3, steps)
    st = wl.kernel_state(g)
    ms_call = time_device(lambda: wl.step_kernel(g, st), 2, steps)           # every kernel of the call, level-major input
    ms_k = time_device.kernel_alone(lambda: wl.step_kernel(g, st), steps)    # the dominant kernel alone
    rec = {'workload': wl.workload, 'columns': int(wl.ncol), 'levels': int(wl.nlev), 'kernel_ms': ms_k, 'call_ms': ms_call,
           'kernel_columns_per_s': wl.ncol / (ms_call * 1e-3), 'device_path_ms_per_step': ms_dev,
           'device_path_columns_per_s': wl.ncol / (ms_dev * 1e-3), 'bytes_per_column': wl.bytes_per_col}
    if wl.kind == 'cape':
        rec['reference_iterations_per_column'] = st['total_iter'] / wl.ncol
    else:
        wl.precision = 'fast'
        time_device(lambda: wl.step_kernel(g, st), 2, 1)
        rec['fast_heights_kernel_ms'] = time_device.kernel_alone(lambda: wl.step_kernel(g, st), steps)
        wl.precision = 'faithful'
    del st
    hp = wl.pinned(d)
    wl.step_e2e(hp, local_rank)
    t0 = time.perf_counter()
    for _ in range(2):
        wl.step_e2e(hp, local_rank)
    torch.cuda.synchronize()
    e2e = (time.perf_counter() - t0) / 2
    h2d, d2h = wl.io_bytes()
    rec['e2e'] = {'value': wl.ncol / e2e, 'ms_per_step': e2e * 1e3, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h}
    del hp, g
    return rec


def strong_scaling_c5(rank, local_rank, world, dev, barrier, max_over_ranks, time_device, args):
    """24 time steps of 721 x 1440 x 137 (most-unstable), column-sharded by time step: 24 / N steps per rank.  Each
    rank holds ONE synthetic step (seed folded with its rank) and runs it 24 / N times — the arithmetic and the bytes
    moved are those of the stack, the host does not spend minutes generating 41 GB of soundings."""
    import torch
    wl = get_workload('C5')
    wl.precision = args.precision
    d = wl.make(rank, 0)
    g = wl.to_device(d, dev)
    mine = 24 // world
    for _ in range(3):
        wl.step_dev(g)
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(mine):
        out = wl.step_dev(g)
    e1.record()
    barrier()
    ms_dev = max_over_ranks(e0.elapsed_time(e1))
    del g, out
    torch.cuda.empty_cache()
    hlm = wl.level_major(d)                          # pinned, in the stack's on-disk order: (time, level, lat, lon)
    for _ in range(2):
        wl.step_e2e(hlm, local_rank, lev_axis=0)
    barrier()
    t0 = time.perf_counter()
    for _ in range(mine):
        wl.step_e2e(hlm, local_rank, lev_axis=0)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    h2d, d2h = wl.io_bytes()
    total = 24 * wl.ncol
    return {'workload': 'calc_cape most-unstable on the 24-step 721x1440x137 stack (configs[4]), 24/N steps per GPU',
            'scaling': 'strong', 'steps_total': 24, 'steps_per_rank': mine, 'columns_total': total, 'levels': wl.nlev,
            'device_resident': {'ms_total': ms_dev, 'value': total / (ms_dev * 1e-3), 'unit': UNIT},
            'e2e': {'ms_total': e2e_s * 1e3, 'value': total / e2e_s, 'unit': UNIT, 'input': 'pinned host numpy, level-major [nlev, ncol]',
                    'h2d_bytes_total': h2d * 24, 'h2d_gb_per_s_per_rank': h2d * mine / e2e_s / 1e9},
            'data': 'each rank re-uses one synthetic 721x1440x137 step for its 24/N steps'}


if __name__ == '__main__':
    main()
