"""End-to-end ms per ERA5 field (pinned host inputs) for several block plans of the host ring."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xcape_b200.cape_cuda import cape
from xcape_b200.synthetic import make_soundings
d = make_soundings('C2', winds=False)
pin = {k: torch.empty(d[k].shape, dtype=torch.float32).pin_memory() for k in ('t', 'td', 'ps', 'ts', 'tds')}
for k in pin: pin[k].numpy()[...] = d[k]
hp = {k: v.numpy() for k, v in pin.items()}
prec = os.environ.get('PREC', 'faithful')
f = lambda: cape(d['p'], hp['t'].T, hp['td'].T, hp['ps'], hp['ts'], hp['tds'], 1, None, 2, 500., 1, 500., 2, precision=prec)
import gc; gc.disable()
for first, chunk, streams in ((1 << 15, 1 << 18, 4), (1 << 16, 1 << 18, 4), (1 << 17, 1 << 18, 4), (1 << 16, 1 << 19, 4),
                              (1 << 15, 1 << 18, 8), (1 << 14, 1 << 18, 8), (1 << 16, 1 << 20, 4), (1 << 17, 1 << 19, 4),
                              (1 << 18, 1 << 18, 4), (1 << 16, 1 << 17, 8)):
    os.environ['XCAPE_B200_FIRST_CHUNK_COLS'] = str(first); os.environ['XCAPE_B200_CHUNK_COLS'] = str(chunk)
    os.environ['XCAPE_B200_STREAMS'] = str(streams)
    for _ in range(3): f()
    ts = []
    for _ in range(12):
        t0 = time.perf_counter(); f(); ts.append(1e3 * (time.perf_counter() - t0))
    ts.sort()
    print(f'{prec} first {first:7d} chunk {chunk:8d} streams {streams}: median {ts[len(ts)//2]:.3f} ms  min {ts[0]:.3f}', flush=True)
