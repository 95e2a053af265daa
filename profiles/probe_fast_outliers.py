"""Which columns does precision='fast' put outside max(1 J/kg, 1e-4 rel), and why?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from xcape_b200 import core
from xcape_b200.synthetic import make_soundings

d = make_soundings('C2')
args = (d['p'], d['t'], d['td'], d['ps'], d['ts'], d['tds'])
kw = dict(source='most-unstable', pinc=500., vertical_lev='pressure', method='cuda')
ex = core.calc_cape(*args, precision='faithful', **kw)
for prec in ('fast', 'fast-relaxed'):
    fa = core.calc_cape(*args, precision=prec, **kw)
    tol = lambda a, r: np.abs(a - r) <= np.maximum(1.0, 1e-4 * np.abs(r))
    bad = np.flatnonzero(~(tol(fa[0], ex[0]) & tol(fa[1], ex[1])))
    print(prec, 'outside:', bad.size, 'zmulev differs:', (fa[3] != ex[3]).sum())
    for c in bad[:12]:
        print(f'  col {c}: cape {ex[0][c]:9.3f} -> {fa[0][c]:9.3f}   cin {ex[1][c]:9.3f} -> {fa[1][c]:9.3f}   z {ex[3][c]:.1f} -> {fa[3][c]:.1f}')
    sub = bad[:64]
    if sub.size:
        o = [oracle.calc_cape_ref(d['p'], d['t'][sub], d['td'][sub], d['ps'][sub], d['ts'][sub], d['tds'][sub], source='most-unstable',
                                  pinc=500., vertical_lev='pressure', tmode=m, contract=c) for m, c in ((0, False), (1, False), (0, True))]
        print('  same columns, oracle LIBM / CR / LIBM+FMA-contraction:')
        for k, c in enumerate(sub[:12]):
            print(f'  col {c}: cape {o[0][0][k]:9.3f} {o[1][0][k]:9.3f} {o[2][0][k]:9.3f}   cin {o[0][1][k]:9.3f} {o[1][1][k]:9.3f} {o[2][1][k]:9.3f}')
