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
    # level axis stored top first (negative level strides / flipped relayout / flipped 1-D pressure axis)
    flip = lambda a: np.ascontiguousarray(a[..., ::-1])
    fa = tuple(flip(d[k]) for k in ('p', 't', 'td', 'u', 'v'))
    surf = (d['ps'], d['ts'], d['tds'])
    for dt in (np.float32, np.float64):
        c = lambda xs: [np.asarray(x, dtype=dt) for x in xs]
        core.calc_cape(*c(fa[:3] + surf), source='most-unstable', vertical_lev=vlev, level_order='top_first')
        lm = [fa[0] if vlev == 'pressure' else np.ascontiguousarray(fa[0].T)] + [np.ascontiguousarray(x.T) for x in fa[1:]]
        core.calc_cape(*c(lm[:3] + list(surf)), source='mixed-layer', vertical_lev=vlev, level_order='auto', lev_axis=0)
        core.calc_srh(*c(list(fa) + list(surf) + [d['us'], d['vs']]), vertical_lev=vlev, level_order='top_first', output_var='all')
        core.calc_srh(*c(lm + list(surf) + [d['us'], d['vs']]), vertical_lev=vlev, level_order='top_first', lev_axis=0)
    core.calc_cape(*(torch.from_numpy(np.ascontiguousarray(x)).cuda() for x in (lm[0], lm[1], lm[2]) + surf), source='surface',
                   vertical_lev=vlev, level_order='top_first', lev_axis=0)
    stdheight(fa[0] if vlev == 'pressure' else fa[0].T, fa[1].T, fa[2].T, *surf, 1 if vlev == 'pressure' else 0, None, 2.,
              2 if vlev == 'pressure' else 1, top_first=True)
    # q -> Td
    from xcape_b200.thermo import dewpoint_from_q
    q = np.full_like(d['t'], 5e-3)
    dewpoint_from_q(d['p'], q)
    dewpoint_from_q(d['p'] if vlev == 'pressure' else np.ascontiguousarray(d['p'].T), np.ascontiguousarray(q.T), lev_axis=0)
    dewpoint_from_q(torch.from_numpy(d['p']).cuda() if vlev == 'pressure' else torch.from_numpy(d['p']).cuda(),
                    torch.from_numpy(q).cuda())
# round 2: sorted execution of the faithful kernel (both orders, forced for this small size), the SRH tile kernel against
# the relayout path, the multi-device entries (one device listed twice)
for cfg, vlev in (('C2', 'pressure'), ('C3', 'sigma')):
    d = make_soundings(cfg, cols=(0, 1500 + 37), active=False)
    args = (d['p'], d['t'], d['td'], d['ps'], d['ts'], d['tds'])
    for mode in ('global', 'window'):
        os.environ.update(XCAPE_B200_SORT='1', XCAPE_B200_SORT_MODE=mode)
        for src in ('surface', 'most-unstable', 'mixed-layer'):
            core.calc_cape(*args, source=src, vertical_lev=vlev, method='cuda')
        core.calc_cape(*args, source='most-unstable', adiabat='reversible-ice', vertical_lev=vlev, method='cuda')
        core.calc_cape(*[torch.from_numpy(a).cuda() for a in args], source='most-unstable', vertical_lev=vlev, method='cuda')
        core.calc_cape(*args, source='most-unstable', vertical_lev=vlev, method='cuda', devices=[0, 0])
    for k in ('XCAPE_B200_SORT', 'XCAPE_B200_SORT_MODE'):
        os.environ.pop(k, None)
    sargs = tuple(d[k] for k in ('p', 't', 'td', 'u', 'v', 'ps', 'ts', 'tds', 'us', 'vs'))
    for tile in ('1', '0'):
        os.environ['XCAPE_B200_SRH_TILE'] = tile
        core.calc_srh(*sargs, vertical_lev=vlev, output_var='all', method='cuda')
        core.calc_srh(*(a.astype(np.float64) for a in sargs), vertical_lev=vlev, output_var='all', method='cuda')
    os.environ.pop('XCAPE_B200_SRH_TILE', None)
    core.calc_srh(*sargs, vertical_lev=vlev, output_var='all', method='cuda', devices=[0, 0])
torch.cuda.synchronize()
print('sanitizer driver done')
