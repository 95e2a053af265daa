"""Wave quantisation of the CAPE kernel: kernel time against the number of columns, in units of one full wave of
resident CTAs (148 SMs x CTAs/SM x columns per CTA).  Linear => no loss; steps => the last partial wave costs a full one.
    python profiles/probe_waves.py [ctas_per_sm] [cols_per_cta]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from xcape_b200.cape_cuda import cape, pres_lev_pos  # noqa: E402
from xcape_b200.synthetic import make_soundings  # noqa: E402

ctas = int(sys.argv[1]) if len(sys.argv) > 1 else 5
cpc = int(sys.argv[2]) if len(sys.argv) > 2 else 256
wave = 148 * ctas * cpc
d = make_soundings('C2', winds=False)
dev = torch.device('cuda', 0)
T = torch.from_numpy(d['t']).to(dev).t().contiguous()
TD = torch.from_numpy(d['td']).to(dev).t().contiguous()
p = torch.from_numpy(d['p']).to(dev)
ps, ts, tds = (torch.from_numpy(d[k]).to(dev) for k in ('ps', 'ts', 'tds'))
for w in (1.0, 2.0, 3.0, 4.0, 4.5, 5.0, 5.25, 5.48):
    n = min(int(w * wave), T.shape[1])
    t, td = T[:, :n].contiguous(), TD[:, :n].contiguous()
    plp = pres_lev_pos(p, ps[:n])
    f = lambda: cape(p, t, td, ps[:n], ts[:n], tds[:n], 1, plp, 2, 500., 1, 500., 2)
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f'{w:5.2f} waves = {n:8d} columns: {ms:7.3f} ms  ({ms / n * 1e6:6.3f} ns/column)', flush=True)
