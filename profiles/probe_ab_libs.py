"""A/B: kernel-only time of the CAPE kernel for several builds of the library (XCAPE_B200_LIB)."""
import os, sys, subprocess, glob
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    sys.path.insert(0, ROOT)
    import torch
    from xcape_b200.cape_cuda import cape, pres_lev_pos
    from xcape_b200.synthetic import make_soundings
    d = make_soundings('C2', winds=False)
    dev = torch.device('cuda', 0)
    t = torch.from_numpy(d['t']).to(dev).t().contiguous(); td = torch.from_numpy(d['td']).to(dev).t().contiguous()
    p = torch.from_numpy(d['p']).to(dev); ps, ts, tds = (torch.from_numpy(d[k]).to(dev) for k in ('ps', 'ts', 'tds'))
    plp = pres_lev_pos(p, ps)
    res = []
    for prec in ('faithful', 'fast'):
        f = lambda: cape(p, t, td, ps, ts, tds, 1, plp, 2, 500., 1, 500., 2, precision=prec)
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): f()
        e1.record(); torch.cuda.synchronize()
        res.append(f'{prec} {e0.elapsed_time(e1)/10:.3f} ms')
    print(os.path.basename(os.environ.get('XCAPE_B200_LIB', 'default')), ' | '.join(res))
    sys.exit(0)
libs = [None] + sorted(glob.glob(os.path.join(ROOT, 'gpurun_scratch', '*.so')))
for lib in libs:
    env = dict(os.environ)
    if lib: env['XCAPE_B200_LIB'] = lib
    print(subprocess.run([sys.executable, __file__, 'child'], env=env, capture_output=True, text=True).stdout.strip())
