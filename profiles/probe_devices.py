"""One process, several GPUs: (1) `devices=[...]` column sharding of one ERA5 field through core.calc_cape,
(2) stream_cape dealing the 24 steps of the C5 stack (721x1440x137, level-major pinned arrays, one synthetic step
re-used 24 times) round-robin to n GPUs.  Prints columns/s for n = 1, 2, 4, 8 (as many as the box has).
    python profiles/probe_devices.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from xcape_b200 import _lib, core, stream  # noqa: E402
from xcape_b200.synthetic import make_soundings  # noqa: E402

ngpu = _lib.device_count()
ns = [n for n in (1, 2, 4, 8) if n <= ngpu]


def pin(a):
    t = torch.empty(a.shape, dtype=torch.float32).pin_memory()
    t.numpy()[...] = a
    return t


d = make_soundings('C2', winds=False)
keep = {k: pin(d[k]) for k in ('t', 'td', 'ps', 'ts', 'tds')}
a = {k: v.numpy() for k, v in keep.items()}
kw = dict(source='most-unstable', pinc=500., vertical_lev='pressure', method='cuda')
ref = None
for n in ns:
    f = lambda: core.calc_cape(d['p'], a['t'], a['td'], a['ps'], a['ts'], a['tds'], devices=list(range(n)), **kw)
    for _ in range(3):
        out = f()
    t0 = time.perf_counter()
    for _ in range(10):
        out = f()
    dt = (time.perf_counter() - t0) / 10
    ref = out if ref is None else ref
    same = all(np.array_equal(x, y) for x, y in zip(out, ref))
    print(f'calc_cape(devices=range({n})) one ERA5 field, pinned reference layout: {dt * 1e3:7.2f} ms, {d["ts"].size / dt:.3e} columns/s, '
          f'identical to 1 GPU: {same}', flush=True)

d5 = make_soundings('C5', grid=(721, 1440), winds=False)
lm = {k: pin(np.ascontiguousarray(d5[k].T)) for k in ('p', 't', 'td')}
sf = {k: pin(d5[k]) for k in ('ps', 'ts', 'tds')}
step = tuple(lm[k].numpy() for k in ('p', 't', 'td')) + tuple(sf[k].numpy() for k in ('ps', 'ts', 'tds'))
skw = dict(source='most-unstable', pinc=500., vertical_lev='sigma', lev_axis=0)
first = None
for n in ns:
    list(stream.stream_cape([step] * n, devices=list(range(n)), **skw))          # warm-up on every GPU
    t0 = time.perf_counter()
    outs = list(stream.stream_cape([step] * 24, devices=list(range(n)), prefetch=2 * n, **skw))
    dt = time.perf_counter() - t0
    first = outs[0] if first is None else first
    same = all(np.array_equal(x, y) for o in outs for x, y in zip(o, first))
    print(f'stream_cape, 24 x 721x1440x137 steps over {n} GPU(s): {dt * 1e3:8.1f} ms total, {24 * d5["ts"].size / dt:.3e} columns/s, '
          f'all steps identical: {same}', flush=True)
