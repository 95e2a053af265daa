"""Disk -> result throughput of the streamed executor on a zarr-v2 archive: writes N synthetic ERA5
pressure-level steps (K / Pa, float32, one chunk per time step and 8 levels) uncompressed and zlib-compressed,
then runs stream_cape over them through xcape_b200.io (page cache warm: this measures decode + staging + GPU,
not the disk of the box).
    python profiles/probe_archive_stream.py [steps] [readers] [levels: 37 or 137]"""
import json
import os
import shutil
import sys
import tempfile
import time
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from xcape_b200 import io as xio, stream  # noqa: E402
from xcape_b200.synthetic import make_soundings  # noqa: E402

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
readers = int(sys.argv[2]) if len(sys.argv) > 2 else 4
nlev = int(sys.argv[3]) if len(sys.argv) > 3 else 37
ny, nx = 721, 1440
cfg = 'C2' if nlev == 37 else 'C5'
d = make_soundings(cfg, winds=False, **({} if nlev == 37 else {'grid': (ny, nx)}))
lm = lambda a: np.ascontiguousarray(a.reshape(ny, nx, nlev).transpose(2, 0, 1))
fields = dict(t=lm(d['t']) + np.float32(273.15), td=lm(d['td']) + np.float32(273.15), sp=d['ps'].reshape(ny, nx) * np.float32(100),
              t2m=d['ts'].reshape(ny, nx) + np.float32(273.15), d2m=d['tds'].reshape(ny, nx) + np.float32(273.15))
if nlev != 37:
    fields['p'] = lm(d['p']) * np.float32(100)
root = tempfile.mkdtemp(prefix='xcape_zarr_')
try:
    for comp in (None, {'id': 'zlib', 'level': 1}):
        base = os.path.join(root, 'zlib' if comp else 'raw')
        nbytes = 0
        for name, a in fields.items():
            path = os.path.join(base, name)
            os.makedirs(path)
            ch = (1, 8, ny, nx) if a.ndim == 3 else (1, ny, nx)
            json.dump(dict(zarr_format=2, shape=[nsteps] + list(a.shape), chunks=list(ch), dtype='<f4', compressor=comp,
                           fill_value=None, order='C', filters=None), open(os.path.join(path, '.zarray'), 'w'))
            for k in range(nsteps):
                if a.ndim == 2:
                    blobs = {f'{k}.0.0': a}
                else:
                    blobs = {}
                    for j in range((nlev + 7) // 8):
                        blk = np.zeros((8, ny, nx), np.float32)
                        blk[:min(8, nlev - 8 * j)] = a[8 * j:8 * j + 8]
                        blobs[f'{k}.{j}.0.0'] = blk
                for fn, blk in blobs.items():
                    raw = blk.tobytes()
                    raw = zlib.compress(raw, 1) if comp else raw
                    nbytes += len(raw)
                    with open(os.path.join(path, fn), 'wb') as f:
                        f.write(raw)
        A = {k: xio.ZarrArray(os.path.join(base, k)) for k in fields}
        kw = dict(source='most-unstable', pinc=500., lev_axis=0)
        if nlev == 37:
            steps = xio.cape_steps(A['t'], A['td'], A['sp'], A['t2m'], A['d2m'], p_levels=d['p'])
            kw['vertical_lev'] = 'pressure'
        else:
            steps = xio.cape_steps(A['t'], A['td'], A['sp'], A['t2m'], A['d2m'], p=A['p'])
            kw['vertical_lev'] = 'sigma'
        list(stream.stream_cape(steps[:2], readers=readers, **kw))          # warm-up (pinned pools, kernels)
        t0 = time.perf_counter()
        out = list(stream.stream_cape(steps, readers=readers, prefetch=max(2, readers), **kw))
        dt = time.perf_counter() - t0
        print(f"{'zlib-1' if comp else 'uncompressed':12s} {nsteps} steps x {ny}x{nx}x{nlev}: {nbytes / 1e9:6.2f} GB on disk, "
              f"{dt / nsteps * 1e3:7.1f} ms/step, {nsteps * ny * nx / dt:.3e} columns/s ({readers} reader threads), "
              f"mean CAPE {float(np.mean(out[-1][0])):.2f}", flush=True)
finally:
    shutil.rmtree(root, ignore_errors=True)
