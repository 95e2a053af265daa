"""Would SORTING the columns by their source parcel remove the SIMT loss of the moist iteration?

Runs the CPU oracle (SPEC arithmetic) with its per-sub-step trace on a strided sample of a synthetic field, builds the
dense pass-count array P[column, absolute layer, sub-step] and evaluates the loop trips a 64-column warp (two
columns per thread, cape_kernel2's policy: absolute-level walk, re-join after every sub-step) executes when the
columns are grouped
  (0) as stored;  (1) sorted by (start level, theta-e of the source parcel);  (2) sorted by theta-e alone;
  (3) sorted by the pass-count vector itself (an upper bound for any key);
against the mean passes per column.
    python profiles/divergence_sorted_model.py [cfg] [ncol] [stride] [shuffle]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from xcape_b200.synthetic import make_soundings, CONFIGS  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else 'C2'
ncol = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
stride = int(sys.argv[3]) if len(sys.argv) > 3 else 8
shuffle = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
source = int(sys.argv[5]) if len(sys.argv) > 5 else 2
ny, nx = CONFIGS[cfg]['grid']
span = min(ncol * stride, ny * nx)
d = make_soundings(cfg, cols=(0, span), winds=False, shuffle=shuffle)
sl = slice(0, span, stride)
p1d = d['p'].ndim == 1
f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
t, td, ps, ts, tds = (f32(d[k][sl]) for k in ('t', 'td', 'ps', 'ts', 'tds'))
p = f32(d['p']) if p1d else f32(d['p'][sl])
ncol, nlev = t.shape
L = oracle.lib()
start = oracle.pres_lev_pos(ps, p[:, None]).astype(np.int32) if p1d else None
cap = 1024 if nlev < 64 else 2048
vp = C.c_void_p
L.xcape_ref_cape_trace.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, C.c_float, C.c_int, C.c_float, C.c_int, vp, C.c_int,
                                   C.c_int64, vp, C.c_int]
from concurrent.futures import ThreadPoolExecutor
tr = np.zeros((ncol, cap), np.int16)
nth = min(16, os.cpu_count() or 4)
def run(i):
    a, b = ncol * i // nth, ncol * (i + 1) // nth
    rc = L.xcape_ref_cape_trace(p.ctypes.data if p1d else p[a:].ctypes.data, t[a:].ctypes.data, td[a:].ctypes.data, int(p1d),
                                ps[a:].ctypes.data, ts[a:].ctypes.data, tds[a:].ctypes.data, 500.0, source, 500.0, 1,
                                start[a:].ctypes.data if p1d else None, nlev, b - a, tr[a:].ctypes.data, cap)
    assert rc == 0
with ThreadPoolExecutor(nth) as ex:
    list(ex.map(run, range(nth)))
assert (tr[:, -1] == 0).all(), 'trace capacity too small'

# dense P[col, abs layer, sub-step]
NS = 64
P = np.zeros((ncol, nlev + 1, NS), np.int16)
first = np.full(ncol, -1)
for i in range(ncol):
    row = tr[i]
    n = int(np.argmax(row == 0)) if (row == 0).any() else cap
    k = -1; s = 0; expect_nloop = False
    for v in row[:n]:
        v = int(v)
        if v <= -1000:
            k = -v - 1000 + (int(start[i]) - 1 if p1d else 0); s = 0; expect_nloop = True
            if first[i] < 0: first[i] = k
        elif v < 0 and expect_nloop:
            expect_nloop = False
        elif v > 0:
            assert s < NS
            P[i, k, s] = v; s += 1
tot = P.sum(axis=(1, 2)).astype(np.int64)
mean = tot.mean()

# source-parcel proxies: MU level (assembled index, 1 = surface) from the oracle, theta-e there (Bolton) in float64
if source == 2:
    out = oracle.calc_cape_ref(d['p'] if p1d else p, t, td, ps, ts, tds, source='most-unstable', pinc=500.,
                               vertical_lev=d['vertical_lev'], tmode=oracle.SPEC, nthreads=nth)
    mulev = np.asarray(out[2]).astype(np.int64).ravel()
else:
    mulev = np.ones(ncol, np.int64)          # surface / mixed-layer: theta-e proxy from the surface values
ks = start.astype(np.int64) if p1d else np.ones(ncol, np.int64)
lev3 = np.clip(ks + mulev - 2, 0, nlev - 1)
idx = np.arange(ncol)
pp = (np.broadcast_to(p, t.shape) if p1d else p).astype(np.float64)
pk = np.where(mulev <= 1, ps, pp[idx, lev3]) * 100.0
tk = np.where(mulev <= 1, ts, t[idx, lev3]).astype(np.float64) + 273.15
tdk = np.where(mulev <= 1, tds, td[idx, lev3]).astype(np.float64) + 273.15
es = 611.2 * np.exp(17.67 * (tdk - 273.15) / (tdk - 29.65))
q = 0.622 * es / (pk - es)
tlcl = 56.0 + 1.0 / (1.0 / (tdk - 56.0) + 0.00125 * np.log(tk / tdk))
the = tk * (1e5 / pk) ** (0.2854 * (1 - 0.28 * q)) * np.exp((3376.0 / tlcl - 2.54) * q * (1 + 0.81 * q))
th = tk * (1e5 / pk) ** 0.2854


def trips(order, W=64):
    n = (len(order) // W) * W
    Q = P[order[:n]].reshape(n // W, W, nlev + 1, NS)
    return Q.max(axis=1).astype(np.int64).sum() / (n // W)


def never(order, W=64):
    n = (len(order) // W) * W
    return tot[order[:n]].reshape(n // W, W).max(axis=1).mean()


print(f'{cfg}: {ncol} columns (every {stride}th of the first {span}), shuffle={shuffle}; mean passes/column {mean:.1f}')
orders = {
    'as stored': np.arange(ncol),
    'sorted by (MU level, theta-e)': np.lexsort((the, mulev)),
    'sorted by theta-e': np.argsort(the, kind='stable'),
    'sorted by (MU level, plcl-ish: theta, theta-e)': np.lexsort((the, np.round(th, 0), mulev)),
    'sorted by (first layer, total passes)  [not realisable]': np.lexsort((tot, first)),
    'sorted by (first layer, theta-e)': np.lexsort((the, first)),
    'sorted by (first layer, ps 2 hPa bins, theta-e)': np.lexsort((the, np.round(ps / 2.0), first)),
    'sorted by (first layer, ps 5 hPa bins, theta-e)': np.lexsort((the, np.round(ps / 5.0), first)),
    'sorted by (first layer, ps 10 hPa bins, theta-e)': np.lexsort((the, np.round(ps / 10.0), first)),
    'sorted by (first layer, theta-e 1 K bins, ps)': np.lexsort((ps, np.round(the), first)),
    'sorted by (first layer, theta-e 2 K bins, ps)': np.lexsort((ps, np.round(the / 2.0), first)),
    'sorted by (first layer, theta-e 4 K bins, ps)': np.lexsort((ps, np.round(the / 4.0), first)),
    'sorted by (first layer, ps)': np.lexsort((ps, first)),
}
for name, o in orders.items():
    print(f'  {name:58s}: {trips(o):8.1f} trips/warp = {trips(o) / mean:.3f} x mean   (never re-join {never(o) / mean:.3f})')

# windowed sorts: the order is sorted inside consecutive windows of W columns only (gathers stay L2-local)
print('windowed (first layer, theta-e bins, ps):')
for W in (8192, 16384, 32768, 65536):
    for tb in (1.0, 2.0, 4.0):
        o = np.concatenate([w0 + np.lexsort((ps[w0:w0 + W], np.round(the[w0:w0 + W] / tb), first[w0:w0 + W]))
                            for w0 in range(0, ncol, W)])
        print(f'  window {W:6d}, theta-e bins {tb:.0f} K : {trips(o) / mean:.3f} x mean')
print('windowed (first layer, theta-e fine):')
for W in (16384, 65536, ncol):
    o = np.concatenate([w0 + np.lexsort((the[w0:w0 + W], first[w0:w0 + W])) for w0 in range(0, ncol, W)])
    print(f'  window {W:6d}: {trips(o) / mean:.3f} x mean')
