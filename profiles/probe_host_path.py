"""Host-pointer path diagnostics: PCIe bandwidth and the effect of block size / stream count."""
import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

if len(sys.argv) > 1 and sys.argv[1] == 'child':
    from xcape_b200.cape_cuda import cape
    from xcape_b200.synthetic import make_soundings
    d = make_soundings('C2', winds=False)
    pin = {k: torch.empty(d[k].shape, dtype=torch.float32).pin_memory() for k in ('t', 'td', 'ps', 'ts', 'tds')}
    for k in pin: pin[k].numpy()[...] = d[k]
    hp = {k: v.numpy() for k, v in pin.items()}
    prec = os.environ.get('PREC', 'faithful')
    f = lambda: cape(d['p'], hp['t'].T, hp['td'].T, hp['ps'], hp['ts'], hp['tds'], 1, None, 2, 500., 1, 500., 2, precision=prec)
    for _ in range(3): f()
    t0 = time.perf_counter()
    for _ in range(10): f()
    dt = (time.perf_counter() - t0) / 10
    # pageable input for comparison
    g = lambda: cape(d['p'], d['t'].T, d['td'].T, d['ps'], d['ts'], d['tds'], 1, None, 2, 500., 1, 500., 2, precision=prec)
    for _ in range(1): g()
    t0 = time.perf_counter()
    for _ in range(3): g()
    dp = (time.perf_counter() - t0) / 3
    print(f"chunk={os.environ.get('XCAPE_B200_CHUNK_COLS')} first={os.environ.get('XCAPE_B200_FIRST_CHUNK_COLS')} streams={os.environ.get('XCAPE_B200_STREAMS')} copy_threads={os.environ.get('XCAPE_B200_COPY_THREADS')}: pinned {dt*1e3:.2f} ms/field ({d['t'].shape[0]/dt/1e6:.1f} Mcol/s)  pageable {dp*1e3:.2f} ms/field")
    sys.exit(0)

x = torch.empty(320 * 1024 * 1024 // 4, dtype=torch.float32).pin_memory()
y = torch.empty_like(x, device='cuda')
for _ in range(2): y.copy_(x, non_blocking=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): y.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f'H2D pinned 320 MiB: {dt*1e3:.2f} ms  ({x.numel()*4/dt/1e9:.1f} GB/s)')
t0 = time.perf_counter()
for _ in range(5): x.copy_(y, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f'D2H pinned 320 MiB: {dt*1e3:.2f} ms  ({x.numel()*4/dt/1e9:.1f} GB/s)')
del x, y
for chunk, first, streams, thr in ((262144, 32768, 4, 8), (262144, 32768, 4, 4), (262144, 32768, 4, 2), (262144, 32768, 4, 16), (131072, 32768, 4, 8), (524288, 65536, 4, 8)):
    env = dict(os.environ, XCAPE_B200_CHUNK_COLS=str(chunk), XCAPE_B200_FIRST_CHUNK_COLS=str(first), XCAPE_B200_STREAMS=str(streams), XCAPE_B200_COPY_THREADS=str(thr))
    print(subprocess.run([sys.executable, __file__, 'child'], env=env, capture_output=True, text=True).stdout.strip())
