"""The second timed step of bench.py's device leg sometimes takes +40..80 ms.  Replicate the exact
sequence and time the host side of every call (torch.empty vs the C call) to see who blocks."""
import gc, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xcape_b200 import _array as A, cape_cuda
from xcape_b200.cape_cuda import cape
from xcape_b200.synthetic import make_soundings

d = make_soundings('C2', winds=False)
dev = torch.device('cuda', 0)
g = {k: torch.from_numpy(d[k]).to(dev) for k in ('p', 't', 'td', 'ps', 'ts', 'tds')}
step = lambda: cape(g['p'], g['t'].t(), g['td'].t(), g['ps'], g['ts'], g['tds'], 1, None, 2, 500., 1, 500., 2)

# time torch.empty inside the shim
orig_empty = A.empty_like_host_or_device
acc = {'empty': 0.0}
def timed_empty(*a, **k):
    t0 = time.perf_counter(); r = orig_empty(*a, **k); acc['empty'] += time.perf_counter() - t0; return r
A.empty_like_host_or_device = timed_empty
cape_cuda.A.empty_like_host_or_device = timed_empty

for trial in range(8):
    for _ in range(3): step()
    torch.cuda.synchronize()
    t_warm = time.time()
    while time.time() - t_warm < 0.5: step()
    torch.cuda.synchronize()
    gc.collect(); gc.disable()
    n = 10
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    host, empt = [], []
    ev[0].record()
    for i in range(n):
        acc['empty'] = 0.0
        t0 = time.perf_counter(); out = step(); host.append(1e3 * (time.perf_counter() - t0)); empt.append(1e3 * acc['empty'])
        ev[i + 1].record()
    torch.cuda.synchronize(); gc.enable()
    devt = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
    print(f'trial {trial}: device ' + ' '.join(f'{x:5.1f}' for x in devt))
    print(f'         host   ' + ' '.join(f'{x:5.1f}' for x in host) + '   torch.empty ' + ' '.join(f'{x:4.1f}' for x in empt), flush=True)
