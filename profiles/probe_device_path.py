"""Where does the device-pointer call spend time beyond the CAPE kernel?  (diagnostic, not a bench)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xcape_b200 import _lib
from xcape_b200.cape_cuda import cape, pres_lev_pos
from xcape_b200.synthetic import make_soundings

d = make_soundings('C2', winds=False)
dev = torch.device('cuda', 0)
g = {k: torch.from_numpy(d[k]).to(dev) for k in ('p', 't', 'td', 'ps', 'ts', 'tds')}
tm, tdm = g['t'].t().contiguous(), g['td'].t().contiguous()
plp = pres_lev_pos(g['p'], g['ps'])

def timeit(name, fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.kernel_launches(); h0 = time.perf_counter(); e0.record()
    for _ in range(n): fn()
    e1.record(); h1 = time.perf_counter(); torch.cuda.synchronize()
    print(f'{name:44s} device {e0.elapsed_time(e1)/n:8.3f} ms/step   host enqueue {1e3*(h1-h0)/n:8.3f} ms/step   launches/step {(_lib.kernel_launches()-l0)/n:.0f}')

A = (1, None, 2, 500., 1, 500., 2)
timeit('level-major + plp given (1 kernel)', lambda: cape(g['p'], tm, tdm, g['ps'], g['ts'], g['tds'], 1, plp, *A[2:]))
timeit('level-major + plp on device (2 kernels)', lambda: cape(g['p'], tm, tdm, g['ps'], g['ts'], g['tds'], *A))
timeit('level-last  + plp given (3 kernels)', lambda: cape(g['p'], g['t'].t(), g['td'].t(), g['ps'], g['ts'], g['tds'], 1, plp, *A[2:]))
timeit('level-last  + plp on device (4 kernels)', lambda: cape(g['p'], g['t'].t(), g['td'].t(), g['ps'], g['ts'], g['tds'], *A))
g64 = {k: v.double() for k, v in g.items()}
timeit('level-last float64 (7 kernels)', lambda: cape(g64['p'], g64['t'].t(), g64['td'].t(), g64['ps'], g64['ts'], g64['tds'], *A))
