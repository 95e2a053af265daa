import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xcape_b200.cape_cuda import cape
from xcape_b200.synthetic import make_soundings
d = make_soundings('C2', winds=False)
pin = {k: torch.empty(d[k].shape, dtype=torch.float32).pin_memory() for k in ('t', 'td', 'ps', 'ts', 'tds')}
for k in pin: pin[k].numpy()[...] = d[k]
hp = {k: v.numpy() for k, v in pin.items()}
f = lambda: cape(d['p'], hp['t'].T, hp['td'].T, hp['ps'], hp['ts'], hp['tds'], 1, None, 2, 500., 1, 500., 2, precision=os.environ.get('PREC', 'faithful'))
for _ in range(3): f()
os.environ['XCAPE_B200_TRACE'] = '1'
for _ in range(2):
    t0 = time.perf_counter(); f(); print(f'python call total {1e3*(time.perf_counter()-t0):.3f} ms', file=sys.stderr)
