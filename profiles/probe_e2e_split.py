"""End-to-end time of one host call when the field is split into k column shards issued concurrently to the
SAME GPU (devices=[0]*k): two rings overlap each other's start-up and tail."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xcape_b200.cape_cuda import cape
from xcape_b200.synthetic import make_soundings
d = make_soundings('C2', winds=False)
pin = {k: torch.empty(d[k].shape, dtype=torch.float32).pin_memory() for k in ('t', 'td', 'ps', 'ts', 'tds')}
for k in pin: pin[k].numpy()[...] = d[k]
hp = {k: v.numpy() for k, v in pin.items()}
for prec in ('faithful', 'fast'):
    for k in (1, 2, 3, 4):
        f = lambda: cape(d['p'], hp['t'].T, hp['td'].T, hp['ps'], hp['ts'], hp['tds'], 1, None, 2, 500., 1, 500., 2,
                         precision=prec, devices=[0] * k)
        for _ in range(3): f()
        t0 = time.perf_counter()
        for _ in range(10): f()
        print(f'{prec:9s} shards on one GPU = {k}: {(time.perf_counter() - t0) * 100:.2f} ms/field', flush=True)
