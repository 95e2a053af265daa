"""Profiling driver: N launches of the CAPE kernel alone (level-major input, start levels
precomputed) on an ERA5-shape block.  Used under ncu; never a bench number.
    python profiles/run_cape_once.py [ncol] [launches] [source] [config] [precision]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from xcape_b200.cape_cuda import cape, pres_lev_pos  # noqa: E402
from xcape_b200.synthetic import make_soundings  # noqa: E402

ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 303104
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
source = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cfg = sys.argv[4] if len(sys.argv) > 4 else 'C2'
prec = sys.argv[5] if len(sys.argv) > 5 else 'faithful'
d = make_soundings(cfg, cols=(0, ncol), winds=False)
dev = torch.device('cuda', 0)
p1d = d['p'].ndim == 1
t = torch.from_numpy(d['t']).to(dev).t().contiguous()
td = torch.from_numpy(d['td']).to(dev).t().contiguous()
p = torch.from_numpy(d['p']).to(dev) if p1d else torch.from_numpy(d['p']).to(dev).t().contiguous()
ps, ts, tds = (torch.from_numpy(d[k]).to(dev) for k in ('ps', 'ts', 'tds'))
plp = pres_lev_pos(p, ps) if p1d else 1
for _ in range(n):
    out = cape(p, t, td, ps, ts, tds, 1 if p1d else 0, plp, source, 500., 1, 500., 2 if p1d else 1, precision=prec)
torch.cuda.synchronize()
print('cape mean', float(out[0].mean()))
