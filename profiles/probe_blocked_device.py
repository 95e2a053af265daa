"""Does cutting a device-resident ERA5 field into the host path's column blocks (one stream per block,
round-robin over 4 streams) cost anything by itself?  Separates the blocking loss from the H2D interplay."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from xcape_b200.cape_cuda import cape
from xcape_b200.synthetic import make_soundings
d = make_soundings('C2', winds=False)
dev = torch.device('cuda', 0)
t = torch.from_numpy(d['t']).to(dev); td = torch.from_numpy(d['td']).to(dev)          # level-last [ncol, nlev]
p = torch.from_numpy(d['p']).to(dev); ps, ts, tds = (torch.from_numpy(d[k]).to(dev) for k in ('ps', 'ts', 'tds'))
n = ps.numel()
plans = {'one call': [n],
         'host-path plan': [32768, 65536, 131072, 262144, 262144, 142336, 142240],
         'equal 4': [n // 4] * 3 + [n - 3 * (n // 4)],
         'equal 8': [n // 8] * 7 + [n - 7 * (n // 8)],
         'fine ramp': [16384] * 4 + [32768] * 2 + [65536] * 2 + [131072, 262144, 262144, n - 4 * 16384 - 2 * 32768 - 2 * 65536 - 131072 - 2 * 262144]}
streams = [torch.cuda.Stream() for _ in range(4)]
def run(plan):
    c0 = 0
    cur = torch.cuda.current_stream()
    ev = torch.cuda.Event(); ev.record(cur)
    for i, m in enumerate(plan):
        s = streams[i % 4] if len(plan) > 1 else cur
        if s is not cur: s.wait_event(ev)
        with torch.cuda.stream(s):
            cape(p, t[c0:c0 + m].T, td[c0:c0 + m].T, ps[c0:c0 + m], ts[c0:c0 + m], tds[c0:c0 + m], 1, None, 2, 500., 1, 500., 2)
        c0 += m
    for s in streams: cur.wait_stream(s)
for name, plan in plans.items():
    assert sum(plan) == n
    for _ in range(3): run(plan)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run(plan)
    e1.record(); torch.cuda.synchronize()
    print(f'{name:16s} {e0.elapsed_time(e1) / 10:7.3f} ms per field', flush=True)
