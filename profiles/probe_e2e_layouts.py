"""End-to-end (host in, host out) time of calc_cape on one ERA5 field for the host-side layouts and staging paths:
pinned level-last (reference layout), pinned level-major with / without the level window, pageable level-last with /
without the window.  XCAPE_B200_TRACE=1 prints the device-side block timeline of one call per case.
    python profiles/probe_e2e_layouts.py [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from xcape_b200 import core  # noqa: E402
from xcape_b200.synthetic import make_soundings  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
d = make_soundings('C2', winds=False)
kw = dict(source='most-unstable', pinc=500., vertical_lev='pressure', method='cuda')


def pin(a):
    t = torch.empty(a.shape, dtype=torch.float32).pin_memory()
    t.numpy()[...] = a
    return t


keep = {k: pin(d[k]) for k in ('t', 'td', 'ps', 'ts', 'tds')}
ll = {k: v.numpy() for k, v in keep.items()}
keep_lm = {k: pin(np.ascontiguousarray(d[k].T)) for k in ('t', 'td')}
lm = dict(ll, **{k: v.numpy() for k, v in keep_lm.items()})
cases = [('pinned level-last (reference layout)', ll, -1, {}),
         ('pinned level-major, level window', lm, 0, {}),
         ('pinned level-major, all levels', lm, 0, {'XCAPE_B200_SHIP_ALL_LEVELS': '1'}),
         ('pageable level-last, level window', d, -1, {}),
         ('pageable level-last, all levels', d, -1, {'XCAPE_B200_SHIP_ALL_LEVELS': '1'})]
for name, a, ax, env in cases:
    os.environ.pop('XCAPE_B200_SHIP_ALL_LEVELS', None)
    os.environ.update(env)
    f = lambda: core.calc_cape(d['p'], a['t'], a['td'], a['ps'], a['ts'], a['tds'], lev_axis=ax, **kw)
    for _ in range(3):
        f()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
    ms = (time.perf_counter() - t0) / reps * 1e3
    print(f'{name:44s} {ms:7.3f} ms per field', flush=True)
    if os.environ.get('XCAPE_B200_TRACE_ONE'):
        os.environ['XCAPE_B200_TRACE'] = '1'
        f()
        del os.environ['XCAPE_B200_TRACE']
