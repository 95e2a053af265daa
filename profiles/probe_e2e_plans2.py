"""Sweep of the host path's block plan (XCAPE_B200_FIRST_CHUNK_COLS / XCAPE_B200_CHUNK_COLS / XCAPE_B200_STREAMS) for one
pinned ERA5 field in the reference layout, through core.calc_cape.  One subprocess per setting (the tunables are read
once per process)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import time
    sys.path.insert(0, ROOT)
    import torch
    from xcape_b200 import core
    from xcape_b200.synthetic import make_soundings
    d = make_soundings('C2', winds=False)
    keep = {}
    for k in ('t', 'td', 'ps', 'ts', 'tds'):
        keep[k] = torch.empty(d[k].shape, dtype=torch.float32).pin_memory()
        keep[k].numpy()[...] = d[k]
    a = {k: v.numpy() for k, v in keep.items()}
    f = lambda: core.calc_cape(d['p'], a['t'], a['td'], a['ps'], a['ts'], a['tds'], source='most-unstable', pinc=500.,
                               vertical_lev='pressure', method='cuda')
    for _ in range(4):
        f()
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for _ in range(10):
            f()
        best = min(best, (time.perf_counter() - t0) / 10)
    print(f'{best * 1e3:.3f}')
    sys.exit(0)
for first in (32768, 65536, 131072):
    for chunk in (131072, 262144, 524288):
        for streams in (3, 4, 6):
            if first > chunk:
                continue
            env = dict(os.environ, XCAPE_B200_FIRST_CHUNK_COLS=str(first), XCAPE_B200_CHUNK_COLS=str(chunk), XCAPE_B200_STREAMS=str(streams))
            r = subprocess.run([sys.executable, __file__, 'child'], env=env, capture_output=True, text=True)
            print(f'first {first:7d} chunk {chunk:7d} streams {streams}: {r.stdout.strip()} ms', flush=True)
