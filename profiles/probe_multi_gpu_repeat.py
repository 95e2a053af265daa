"""devices=[0, 1, ...] called repeatedly: per-call time must stay flat (no per-call stream / buffer leak)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from xcape_b200 import _lib, core
from xcape_b200.synthetic import make_soundings
n = min(_lib.device_count(), 8)
d = make_soundings('C2', winds=False)
args = (d['p'], d['t'], d['td'], d['ps'], d['ts'], d['tds'])
kw = dict(source='most-unstable', pinc=500., vertical_lev='pressure', method='cuda', devices=list(range(n)))
ts = []
for i in range(40):
    t0 = time.perf_counter(); out = core.calc_cape(*args, **kw); ts.append(1e3 * (time.perf_counter() - t0))
print(f'{n} GPUs, pageable numpy inputs: first 5 calls {np.round(ts[:5], 1)}, calls 6-20 mean {np.mean(ts[5:20]):.1f} ms, last 20 mean {np.mean(ts[20:]):.1f} ms')
one = core.calc_cape(*args, **{**kw, 'devices': [0]})
print('equal to single GPU:', all(np.array_equal(a, b) for a, b in zip(out, one)))
