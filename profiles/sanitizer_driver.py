"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / initcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from xcape_b200 import core
from xcape_b200.stdheight_cuda import stdheight
from xcape_b200.srh_cuda import srh as srh_two_call
from xcape_b200.synthetic import make_soundings

for cfg, vlev in (('C2', 'pressure'), ('C3', 'sigma')):
    d = make_soundings(cfg, cols=(0, 1500 + 37), active=False)
    args = (d['p'], d['t'], d['td'], d['ps'], d['ts'], d['tds'])
    for src in ('surface', 'most-unstable', 'mixed-layer'):
        for prec in ('faithful', 'fast', 'fast-relaxed'):
            core.calc_cape(*args, source=src, vertical_lev=vlev, method='cuda', precision=prec)
    core.calc_cape(*(a.astype(np.float64) for a in args), source='most-unstable', adiabat='reversible-ice', vertical_lev=vlev, method='cuda')
    dev = [torch.from_numpy(a).cuda() for a in args]
    core.calc_cape(*dev, source='most-unstable', vertical_lev=vlev, method='cuda')
    sargs = tuple(d[k] for k in ('p', 't', 'td', 'u', 'v', 'ps', 'ts', 'tds', 'us', 'vs'))
    p = d['p'].copy()
    if vlev == 'sigma':
        p[::3, 7] = p[::3, 6]                          # EXACT path through the work list
        sargs = (p,) + sargs[1:]
    for prec in ('faithful', 'fast'):
        core.calc_srh(*sargs, vertical_lev=vlev, output_var='all', method='cuda', precision=prec)
    core.calc_srh(*(torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in sargs), vertical_lev=vlev, method='cuda')
    p2 = d['p'] if vlev == 'pressure' else d['p'].T
    H, Hs = stdheight(p2, d['t'].T, d['td'].T, d['ps'], d['ts'], d['tds'], 1 if vlev == 'pressure' else 0, None if vlev == 'pressure' else 1, 2., 2 if vlev == 'pressure' else 1)
    srh_two_call(d['u'].T, d['v'].T, H, d['us'], d['vs'], Hs, 1, 3000, 1, 2)
torch.cuda.synchronize()
print('sanitizer driver done')
