"""Why does the device-path leg sometimes run slower than the kernel alone?  Per-step event timing,
with and without an nvidia-smi poller."""
import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from xcape_b200.cape_cuda import cape
from xcape_b200.synthetic import make_soundings

d = make_soundings('C2', winds=False)
dev = torch.device('cuda', 0)
g = {k: torch.from_numpy(d[k]).to(dev) for k in ('p', 't', 'td', 'ps', 'ts', 'tds')}
step = lambda: cape(g['p'], g['t'].t(), g['td'].t(), g['ps'], g['ts'], g['tds'], 1, None, 2, 500., 1, 500., 2)

def rep(tag, n=10):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(n):
        step(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
    print(f'{tag:28s} mean {sum(ts)/n:7.2f} ms  per-step: ' + ' '.join(f'{t:5.1f}' for t in ts), flush=True)

for i in range(3): rep(f'cold rep {i}')
time.sleep(1.0)
for i in range(2): rep(f'after 1 s idle rep {i}')
p = subprocess.Popen(['nvidia-smi', '--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active', '--format=csv,noheader', '-lms', '100'],
                     stdout=subprocess.PIPE, text=True)
time.sleep(0.05)
for i in range(3): rep(f'nvidia-smi starting rep {i}')
time.sleep(2.0)
for i in range(3): rep(f'nvidia-smi steady rep {i}')
p.terminate()
print(p.stdout.read()[-600:])
for i in range(2): rep(f'poller gone rep {i}')

# --- does the stream-ordered pool reuse blocks across steps that are enqueued but not yet executed?
torch.cuda.synchronize()
free0 = torch.cuda.mem_get_info()[0]
t0 = time.perf_counter()
for i in range(300):
    step()
    if i in (0, 9, 49, 99, 199, 299):
        print(f'enqueued {i+1:3d} steps: host {time.perf_counter()-t0:6.2f} s, device memory in use grew by {(free0 - torch.cuda.mem_get_info()[0])/1e9:6.2f} GB', flush=True)
torch.cuda.synchronize()
print(f'after sync: grew by {(free0 - torch.cuda.mem_get_info()[0])/1e9:6.2f} GB')
for i in range(2): rep(f'after deep queue rep {i}')
