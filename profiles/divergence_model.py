"""What does SIMT cost the moist iteration?  Runs the CPU oracle (SPEC arithmetic) with its per-sub-step trace on
consecutive synthetic columns, groups them 32 to a warp as the kernel does, and compares the loop trips a warp
executes under three synchronisation policies:
  (a) lanes re-join after every sub-step  (the shipping kernel):  sum over sub-steps of max over lanes
  (b) lanes re-join after every layer:                             sum over layers of max over lanes of the layer's passes
  (c) lanes never re-join:                                          max over lanes of the column's passes
against the mean passes per column (the work a perfectly packed machine would do).
    python profiles/divergence_model.py [cfg] [ncol] [source] [offset]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from xcape_b200.synthetic import make_soundings  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else 'C2'
ncol = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
source = int(sys.argv[3]) if len(sys.argv) > 3 else 2
off = int(sys.argv[4]) if len(sys.argv) > 4 else 300000
d = make_soundings(cfg, cols=(off, off + ncol), winds=False)
p1d = d['p'].ndim == 1
L = oracle.lib()
f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
t, td, ps, ts, tds, p = (f32(d[k]) for k in ('t', 'td', 'ps', 'ts', 'tds', 'p'))
nlev = t.shape[1]
start = oracle.pres_lev_pos(ps, p[:, None]).astype(np.int32) if p1d else None
cap = 4096
tr = np.zeros((ncol, cap), np.int16)
vp = C.c_void_p
L.xcape_ref_cape_trace.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, C.c_float, C.c_int, C.c_float, C.c_int, vp, C.c_int,
                                   C.c_int64, vp, C.c_int]
rc = L.xcape_ref_cape_trace(p.ctypes.data, t.ctypes.data, td.ctypes.data, int(p1d), ps.ctypes.data, ts.ctypes.data,
                            tds.ctypes.data, 500.0, source, 500.0, 1, start.ctypes.data if p1d else None, nlev, ncol,
                            tr.ctypes.data, cap)
assert rc == 0
# parse: per column list of layers, each a list of pass counts
cols, klev = [], []
for i in range(ncol):
    row = tr[i]
    layers, levels = [], []
    for v in row:
        if v == 0:
            break
        if v <= -1000:
            levels.append(-int(v) - 1000 + (int(start[i]) - 1 if p1d else 0))   # absolute level index of the layer top
            layers.append([])
        elif v > 0:
            layers[-1].append(int(v))
    cols.append(layers)
    klev.append(levels)
tot_mean = np.mean([sum(sum(l) for l in c) for c in cols])
a_trips = b_trips = c_trips = 0.0
a_sub = 0
nw = ncol // 32
for w in range(nw):
    lanes = cols[w * 32:(w + 1) * 32]
    nl = max(len(c) for c in lanes)
    c_trips += max(sum(sum(l) for l in c) for c in lanes)
    for k in range(nl):
        # layers are aligned from the END of the ascent for lanes that start higher (MU source): align by index from
        # each lane's own start instead — the kernel walks lanes in lock-step by their own layer counter
        lay = [c[k] if k < len(c) else [] for c in lanes]
        b_trips += max(sum(l) for l in lay)
        ns = max(len(l) for l in lay)
        a_sub += ns
        for n in range(ns):
            a_trips += max(l[n] if n < len(l) else 0 for l in lay)
# (d) lanes walk the layers by ABSOLUTE level (a lane whose parcel starts higher waits for the warp to get there),
#     re-joining after every sub-step; (e) same, re-joining after every layer
d_trips = e_trips = 0.0
for w in range(nw):
    lanes = list(zip(cols[w * 32:(w + 1) * 32], klev[w * 32:(w + 1) * 32]))
    ks = sorted({k for _, kl in lanes for k in kl})
    for k in ks:
        lay = [c[kl.index(k)] if k in kl else [] for c, kl in lanes]
        e_trips += max(sum(l) for l in lay)
        for n in range(max(len(l) for l in lay)):
            d_trips += max(l[n] if n < len(l) else 0 for l in lay)
print(f'{cfg} source={source}: {ncol} columns from {off}; mean passes/column {tot_mean:.1f}; sub-steps per warp {a_sub / nw:.1f}')
print(f'  (a) re-join every sub-step : {a_trips / nw:8.1f} trips/warp  = {a_trips / nw / tot_mean:.3f} x mean')
print(f'  (b) re-join every layer    : {b_trips / nw:8.1f} trips/warp  = {b_trips / nw / tot_mean:.3f} x mean')
print(f'  (d) by absolute level, re-join every sub-step: {d_trips / nw:8.1f} = {d_trips / nw / tot_mean:.3f} x mean')
print(f'  (e) by absolute level, re-join every layer   : {e_trips / nw:8.1f} = {e_trips / nw / tot_mean:.3f} x mean')
print(f'  (c) never re-join          : {c_trips / nw:8.1f} trips/warp  = {c_trips / nw / tot_mean:.3f} x mean')
