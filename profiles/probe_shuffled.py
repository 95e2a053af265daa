"""Kernel time on spatially smooth vs iid-shuffled columns (the adversarial case for warp divergence)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xcape_b200.cape_cuda import cape, pres_lev_pos
from xcape_b200.synthetic import make_soundings
dev = torch.device('cuda', 0)
for shuffle in (False, True):
    for active in (True, False):
        d = make_soundings('C2', winds=False, shuffle=shuffle, active=active)
        t = torch.from_numpy(d['t']).to(dev).t().contiguous(); td = torch.from_numpy(d['td']).to(dev).t().contiguous()
        p = torch.from_numpy(d['p']).to(dev); ps, ts, tds = (torch.from_numpy(d[k]).to(dev) for k in ('ps', 'ts', 'tds'))
        plp = pres_lev_pos(p, ps)
        for prec in ('faithful', 'fast'):
            f = lambda: cape(p, t, td, ps, ts, tds, 1, plp, 2, 500., 1, 500., 2, precision=prec, return_counters=True)
            for _ in range(3): out = f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): f()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            it = out[5].double().sum().item()
            print(f'shuffle={shuffle!s:5} active={active!s:5} {prec:9s}: {ms:7.3f} ms/field, {it / ms * 1e-6:8.1f} G passes/s, {it / 1038240:7.1f} passes/column', flush=True)
