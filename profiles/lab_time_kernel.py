"""Lab driver: time the CAPE kernel alone (level-major, device-resident, start levels precomputed) on a full
synthetic field and print ms per launch plus output checksums, for A/B-ing builds of the library:
    XCAPE_B200_LIB=/path/to/variant.so python profiles/lab_time_kernel.py [cfg] [source] [launches] [precision]
Not a bench number (no clocks sampling); the checksums let two variants be compared bit for bit."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from xcape_b200 import _lib  # noqa: E402
from xcape_b200.cape_cuda import cape, pres_lev_pos  # noqa: E402
from xcape_b200.synthetic import make_soundings  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else 'C2'
source = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = int(sys.argv[3]) if len(sys.argv) > 3 else 10
prec = sys.argv[4] if len(sys.argv) > 4 else 'faithful'
kw = {'grid': (721, 1440)} if cfg == 'C5' else {}
d = make_soundings(cfg, winds=False, shuffle=bool(int(os.environ.get('LAB_SHUFFLE', '0'))), active=bool(int(os.environ.get('LAB_ACTIVE', '1'))), **kw)
dev = torch.device('cuda', 0)
p1d = d['p'].ndim == 1
t = torch.from_numpy(d['t']).to(dev).t().contiguous()
td = torch.from_numpy(d['td']).to(dev).t().contiguous()
p = torch.from_numpy(d['p']).to(dev) if p1d else torch.from_numpy(d['p']).to(dev).t().contiguous()
ps, ts, tds = (torch.from_numpy(d[k]).to(dev) for k in ('ps', 'ts', 'tds'))
plp = pres_lev_pos(p, ps) if p1d else 1


def run(**k):
    return cape(p, t, td, ps, ts, tds, 1 if p1d else 0, plp, source, 500., 1, 500., 2 if p1d else 1, precision=prec, **k)


for _ in range(3):
    out = run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    out = run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
out = run(return_counters=True)
h = hashlib.sha1(b''.join(o.cpu().numpy().tobytes() for o in out)).hexdigest()[:16]
print(f'{os.path.basename(_lib.LIB_PATH)} {cfg} source={source} {prec} sort={os.environ.get("XCAPE_B200_SORT", "-")} shuffle={os.environ.get("LAB_SHUFFLE", "0")} active={os.environ.get("LAB_ACTIVE", "1")}: {ms:.3f} ms/launch, {t.shape[1] / ms * 1e-3:.3e} col/s, '
      f'iters/col {float(out[5].double().mean()):.1f}, cape mean {float(out[0].double().mean()):.4f}, sha1 {h}', flush=True)
