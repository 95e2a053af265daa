"""Profiling driver: N launches of the fused SRH kernel (level-major device-resident HRRR-shape input).
Used under ncu; never a bench number.    python profiles/run_srh_once.py [ncol] [launches] [precision]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from xcape_b200.srh_cuda import srh_fused  # noqa: E402
from xcape_b200.synthetic import make_soundings  # noqa: E402
ncol = int(sys.argv[1]) if len(sys.argv) > 1 else 1905141
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prec = sys.argv[3] if len(sys.argv) > 3 else 'faithful'
d = make_soundings('C4', cols=(0, ncol), winds=True)
dev = torch.device('cuda', 0)
f3 = [torch.from_numpy(d[k]).to(dev).t().contiguous() for k in ('p', 't', 'td', 'u', 'v')]
f1 = [torch.from_numpy(d[k]).to(dev) for k in ('ps', 'ts', 'tds', 'us', 'vs')]
for _ in range(n):
    out = srh_fused(*[x for x in f3], *f1, 0, None, 3000., 2., 1, 1, precision=prec)
torch.cuda.synchronize()
print('srh_rm mean', float(out[0].mean()))
