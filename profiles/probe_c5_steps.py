"""Per-step device times of 24 consecutive C5 steps through the device path in the reference layout (bench's strong-scaling leg)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
wl = bench.get_workload('C5')
wl.precision = 'faithful'
d = wl.make(0, 0)
g = wl.to_device(d, torch.device('cuda', 0))
wl.step_dev(g)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(25)]
ev[0].record()
for i in range(24):
    out = wl.step_dev(g)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [round(a.elapsed_time(b), 2) for a, b in zip(ev[:-1], ev[1:])]
print('DIRECT', os.environ.get('XCAPE_B200_CAPE_DIRECT', '1'), 'total', round(sum(ms), 1), ms)
