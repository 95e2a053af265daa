"""Streamed multi-time-step execution (xcape_b200.stream): N ERA5-sized time steps held as level-major
.npy files (memory-mapped; /dev/shm so that the 'disk' is page cache) against the same steps as direct
calls on arrays already in host memory."""
import os, sys, time, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from xcape_b200 import core
from xcape_b200.stream import stream_cape
from xcape_b200.synthetic import make_soundings

nstep = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prec = sys.argv[2] if len(sys.argv) > 2 else 'faithful'
d = make_soundings('C2', winds=False)
root = tempfile.mkdtemp(dir='/dev/shm')
try:
    tl, tdl = np.ascontiguousarray(d['t'].T), np.ascontiguousarray(d['td'].T)
    steps, arrays = [], []
    for k in range(nstep):
        np.save(f'{root}/t{k}.npy', tl); np.save(f'{root}/td{k}.npy', tdl)
        steps.append((d['p'], np.load(f'{root}/t{k}.npy', mmap_mode='r'), np.load(f'{root}/td{k}.npy', mmap_mode='r'),
                      d['ps'], d['ts'], d['tds']))
        arrays.append((d['p'], tl, tdl, d['ps'], d['ts'], d['tds']))
    kw = dict(source='most-unstable', vertical_lev='pressure', lev_axis=0, precision=prec)
    core.calc_cape(*arrays[0], **kw)
    for name, run in (('direct calls, arrays in memory', lambda: [core.calc_cape(*a, **kw) for a in arrays]),
                      ('stream unpinned, 2 workers 2 readers', lambda: list(stream_cape(steps, devices=[0, 0], readers=2, prefetch=3, pinned=False, **kw))),
                      ('stream, 1 worker 1 reader', lambda: list(stream_cape(steps, **kw))),
                      ('stream, 1 worker 2 readers', lambda: list(stream_cape(steps, readers=2, prefetch=3, **kw))),
                      ('stream, 2 workers 2 readers', lambda: list(stream_cape(steps, devices=[0, 0], readers=2, prefetch=3, **kw))),
                      ('stream, 2 workers 4 readers', lambda: list(stream_cape(steps, devices=[0, 0], readers=4, prefetch=4, **kw)))):
        run()
        t0 = time.perf_counter(); run(); dt = time.perf_counter() - t0
        print(f'{prec:9s} {name:38s} {1e3 * dt / nstep:7.2f} ms/step  {nstep * d["ps"].size / dt:.3e} col/s', flush=True)
finally:
    shutil.rmtree(root)
