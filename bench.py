#!/usr/bin/env python
"""bench.py — CAPE/CIN columns/sec on the ERA5-shape most-unstable workload (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on host cores
    (N > 1: launched by torchrun, one rank per GPU)

One "step" = one pass of the hot path (calc_cape, most-unstable parcel, pinc = 500 Pa,
pseudo-liquid adiabat) over one synthetic ERA5 pressure-level field: 721 x 1440 columns x 37
levels (BASELINE configs[1], "C2").  Weak scaling: every rank owns one such field (a different
time step of the stack — SURVEY §8e shards time chunks / column blocks, no collective).

Numbers on the JSON line:
  value      whole-job columns/s with the field resident in HBM in the reference's own layout
             ([ncol, nlev] float32, level last); the timed call is the public device-pointer
             path: pres_lev_pos + relayout + CAPE kernel, CUDA events on the launch stream.
  e2e        same metric through the host-buffer C-ABI call (pinned host numpy in, host numpy
             out, H2D / D2H inside the timed region).
  roofline   the CAPE kernel alone (level-major input, one launch per step, CUDA events):
             algorithmic work = 73 flop x (moist iterations executed, counted by the kernel
             and asserted equal to the oracle's count in tests) against the FFMA peak measured
             on this GPU by xcape_cuda_measure_peaks; plus the HBM view (324 B/column).
  cpu_baseline  the CPU oracle in LIBM mode (= the reference algorithm with the libm gfortran
             links) on all host threads over a bounded sample of the same field (rank 0, N=1).
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

import numpy as np  # noqa: E402

METRIC = 'CAPE/CIN columns/sec (ERA5-shape, MU parcel)'
UNIT = 'columns/s'
WORKLOAD = 'calc_cape most-unstable, ERA5 pressure levels 721x1440x37, pinc=500 Pa, pseudo-liquid (configs[1])'
FLOP_PER_ITER = 73.0
BYTES_PER_COL = 324.0         # 4*(2*37) + 12 + 16


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


def cpu_reference_rate(d, ncol_sample, nthreads, counters=False):
    """Reference algorithm (oracle, LIBM arithmetic) on `ncol_sample` columns spread evenly over
    the field; returns (columns/s, mean iterations/column or None)."""
    import oracle
    ncol = d['t'].shape[0]
    stride = max(1, ncol // ncol_sample)
    idx = np.arange(0, ncol, stride)[:ncol_sample]
    sub = [np.ascontiguousarray(d[k][idx]) for k in ('t', 'td', 'ps', 'ts', 'tds')]
    t0 = time.perf_counter()
    out = oracle.calc_cape_ref(d['p'], *sub, source='most-unstable', pinc=500., adiabat='pseudo-liquid',
                               vertical_lev='pressure', tmode=oracle.LIBM, nthreads=nthreads, counters=counters)
    dt = time.perf_counter() - t0
    it = float(out[1]['n_iter'].mean()) if counters else None
    return len(idx) / dt, it, len(idx)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm on this box's host cores."""
    if rank != 0:
        return
    import oracle
    from xcape_b200.synthetic import make_soundings
    oracle.build()
    nth = host_threads()
    d = make_soundings('C2', cols=(0, args.cols) if args.cols else None, winds=False)
    ncol = d['t'].shape[0]
    rate0, _, _ = cpu_reference_rate(d, min(ncol, 4000 * nth), nth)
    sample = int(min(ncol, max(2000 * nth, rate0 * args.ref_seconds)))
    for _ in range(args.warmup):
        cpu_reference_rate(d, sample, nth)
    t0 = time.perf_counter()
    n_done = 0
    for _ in range(args.steps):
        _, _, n = cpu_reference_rate(d, sample, nth)
        n_done += n
    dt = time.perf_counter() - t0
    v = n_done / dt
    desc = f'{sample} of {ncol} columns per step (evenly strided), oracle tmode=LIBM, {nth} threads'
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'sample': desc},
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': nth, 'kind': 'port', 'sample': desc},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='cuda', choices=['cuda', 'reference'])
    ap.add_argument('--cols', type=int, default=0, help='debug: use only the first COLS columns of the field')
    ap.add_argument('--cpu-seconds', type=float, default=12.0, help='CPU work budget of the cpu_baseline leg')
    ap.add_argument('--ref-seconds', type=float, default=3.0, help='CPU seconds per step of --impl reference')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'cuda' else args.warmup

    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))

    if args.impl == 'reference':
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from xcape_b200 import _lib
    from xcape_b200.cape_cuda import cape as cape_cuda, pres_lev_pos
    from xcape_b200.synthetic import make_soundings

    if not torch.cuda.is_available() or _lib.device_count() < 1:
        raise SystemExit('bench.py: no CUDA device — the CUDA path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
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
    d = make_soundings('C2', seed=2 + rank, cols=(0, args.cols) if args.cols else None, winds=False)
    ncol, nlev = d['t'].shape
    common = dict(flag_1d=1, source=2, ml_depth=500., adiabat=1, pinc=500., type_grid=2)

    def run_cape(p, t2d, td2d, ps, ts, tds, plp=None, **kw):
        return cape_cuda(p, t2d, td2d, ps, ts, tds, common['flag_1d'], plp, common['source'], common['ml_depth'],
                         common['adiabat'], common['pinc'], common['type_grid'], **kw)

    # device-resident copies, reference layout ([ncol, nlev], level last)
    g = {k: torch.from_numpy(d[k]).to(dev) for k in ('p', 't', 'td', 'ps', 'ts', 'tds')}
    sampler = ClockSampler(local_rank)
    if rank == 0:
        # nvidia-smi's start-up (NVML init) stalls the GPU for tens of ms: let it reach its steady
        # sampling state before anything is timed
        sampler.start()
        t_wait = time.time()
        while not sampler.rows and time.time() - t_wait < 5.0:
            time.sleep(0.05)
    windows = []

    def step_dev():
        return run_cape(g['p'], g['t'].t(), g['td'].t(), g['ps'], g['ts'], g['tds'])

    for _ in range(args.warmup):
        step_dev()
    barrier()
    t_warm = time.time()
    while time.time() - t_warm < 0.5:      # clocks / power state settled (untimed)
        step_dev()
    barrier()
    l0 = _lib.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        out_dev = step_dev()
    e1.record()
    barrier()
    windows.append((w0, time.time()))
    launches = _lib.kernel_launches() - l0
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    value = world * ncol / (ms_step * 1e-3)

    # ---- roofline leg: the CAPE kernel alone (level-major input, start precomputed) ---------
    tm = g['t'].t().contiguous()
    tdm = g['td'].t().contiguous()
    plp = pres_lev_pos(g['p'], g['ps'])
    cnt = run_cape(g['p'], tm, tdm, g['ps'], g['ts'], g['tds'], plp, return_counters=True)
    total_iter = float(cnt[5].to(torch.float64).sum().item())
    for _ in range(3):
        run_cape(g['p'], tm, tdm, g['ps'], g['ts'], g['tds'], plp)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lk = _lib.kernel_launches()
    w0 = time.time()
    k0.record()
    for _ in range(args.steps):
        run_cape(g['p'], tm, tdm, g['ps'], g['ts'], g['tds'], plp)
    k1.record()
    torch.cuda.synchronize()
    windows.append((w0, time.time()))
    assert _lib.kernel_launches() - lk == args.steps, 'roofline leg must be exactly one kernel per step'
    ms_kernel = k0.elapsed_time(k1) / args.steps
    del tm, tdm

    # ---- end to end: pinned host buffers through the host-pointer C-ABI call ------------------
    pin = {k: torch.empty(d[k].shape, dtype=torch.float32).pin_memory() for k in ('t', 'td', 'ps', 'ts', 'tds')}
    for k in pin:
        pin[k].numpy()[...] = d[k]
    hp = {k: v.numpy() for k, v in pin.items()}

    def step_e2e():
        return run_cape(d['p'], hp['t'].T, hp['td'].T, hp['ps'], hp['ts'], hp['tds'], device=local_rank)

    for _ in range(2):
        out_host = step_e2e()
    barrier()
    w0 = time.time()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_host = step_e2e()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    barrier()
    windows.append((w0, time.time()))
    e2e_value = world * ncol / e2e_s
    h2d = int(2 * ncol * nlev * 4 + 3 * ncol * 4 + nlev * 4)
    d2h = int(16 * ncol)
    same = all(np.array_equal(a, b.cpu().numpy()) for a, b in zip(out_host, out_dev))

    clocks = sampler.stop(windows) if rank == 0 else None

    # ---- peaks + CPU baseline (rank 0) ---------------------------------------------------------
    if rank == 0:
        fp32_peak, fp64_peak = _lib.measure_peaks(local_rank, 5)
        peaks_file = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        hbm_peak, hbm_src = 6650.0, 'fallback (B200_PROFILING.md)'
        if os.path.exists(peaks_file):
            hbm_peak, hbm_src = float(json.load(open(peaks_file))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        achieved_tf = FLOP_PER_ITER * total_iter / (ms_kernel * 1e-3) / 1e12
        achieved_gbs = BYTES_PER_COL * ncol / (ms_kernel * 1e-3) / 1e9
        roofline = {
            'bound': 'fp32', 'achieved': achieved_tf, 'peak': fp32_peak, 'unit': 'TFLOP/s', 'frac': achieved_tf / fp32_peak,
            'traffic': None, 'kernel': 'cape_kernel<MathSpec,2,1,true>', 'kernel_ms': ms_kernel,
            'work': f'{FLOP_PER_ITER:.0f} flop x {total_iter / ncol:.1f} moist iterations/column (counted by the kernel)',
            'peak_source': 'FFMA microbenchmark on this GPU (xcape_cuda_measure_peaks), 2 flop/FMA',
            'fp64_peak_tflops': fp64_peak, 'iterations_per_s': total_iter / (ms_kernel * 1e-3),
            'hbm': {'bound': 'hbm', 'achieved': achieved_gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': achieved_gbs / hbm_peak,
                    'bytes_per_column': BYTES_PER_COL, 'peak_source': hbm_src}}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            import oracle
            oracle.build()
            nth = host_threads()
            r0, _, _ = cpu_reference_rate(d, min(ncol, 2000 * nth), nth)
            sample = int(min(ncol, max(1000 * nth, r0 * args.cpu_seconds)))
            rate, it_cpu, n = cpu_reference_rate(d, sample, nth, counters=True)
            cpu = {'value': rate, 'unit': UNIT, 'cores': nth, 'kind': 'port',
                   'sample': f'{n} of {ncol} columns (evenly strided) of the same field, oracle tmode=LIBM '
                             f'(reference algorithm, glibc libm), {nth} threads; {it_cpu:.1f} iterations/column'}
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'columns_per_gpu': ncol, 'levels': nlev, 'layout': 'level-last [ncol, nlev] float32 '
                       '(reference layout), resident in HBM', 'precision': 'faithful (bit-exact vs oracle SPEC arithmetic)',
                       'l2': f'inputs {2 * ncol * nlev * 4 / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)',
                       'parallelism': f'{world} x independent column shards, no collective'},
            'clocks': clocks, 'gpu_launches': int(launches),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': e2e_s * 1e3, 'matches_device_path': bool(same)},
            'roofline': roofline, 'cpu_baseline': cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
