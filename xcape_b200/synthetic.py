"""Deterministic synthetic soundings for parity tests and benchmarks.

Every column is a function of ``(seed, global column index)`` only: random draws come from a
counter-based Philox generator keyed per block of ``BLOCK`` columns, so any shard of any
configuration can be regenerated identically on any rank without generating the rest.

Per-column scalars (surface pressure / temperature / dew-point depression, lapse rate,
tropopause temperature) are smooth low-wavenumber fields over the (ny, nx) grid plus a little
iid noise, so that neighbouring columns resemble each other as in real analyses; ``shuffle=True``
permutes the columns (the adversarial case for warp divergence).
"""
import numpy as np

BLOCK = 4096

ERA5_LEVELS_HPA = np.array(
    [1000, 975, 950, 925, 900, 875, 850, 825, 800, 775, 750, 700, 650, 600, 550, 500, 450, 400, 350,
     300, 250, 225, 200, 175, 150, 125, 100, 70, 50, 30, 20, 10, 7, 5, 3, 2, 1], dtype=np.float32)

# name -> (grid (ny, nx), nlev, vertical_lev, seed)   — BASELINE.json configs, SURVEY §8 sizes
CONFIGS = {
    'C1': dict(grid=(25, 40), nlev=50, vertical_lev='sigma', seed=1),          # 1000 columns x 50
    'C2': dict(grid=(721, 1440), nlev=37, vertical_lev='pressure', seed=2),    # ERA5 pressure levels
    'C3': dict(grid=(1059, 1799), nlev=50, vertical_lev='sigma', seed=3),      # HRRR model levels
    'C4': dict(grid=(1059, 1799), nlev=50, vertical_lev='sigma', seed=3),      # HRRR, SRH
    'C5': dict(grid=(24 * 721, 1440), nlev=137, vertical_lev='sigma', seed=5),  # 24 h x ERA5 x 137
}


def sigma_levels(nlev):
    """Fixed analytic sigma list: linear 1 -> 0.02 for <= 64 levels, otherwise a stretched
    list reaching 1e-4 at the top (stand-in for the ECMWF L137 a/b table)."""
    if nlev <= 64:
        return np.linspace(1.0, 0.02, nlev).astype(np.float64)
    x = np.linspace(0.0, 1.0, nlev)
    return np.exp(np.log(1e-4) * x ** 1.6)


def _smooth01(ci, grid, k1, k2, ph, noise):
    ny, nx = grid
    i = (ci // nx).astype(np.float64) / max(ny, 1)
    j = (ci % nx).astype(np.float64) / max(nx, 1)
    f = 0.5 + 0.35 * np.sin(2 * np.pi * (k1 * i + ph)) * np.cos(2 * np.pi * (k2 * j + 0.37 * ph)) \
        + 0.15 * (noise - 0.5)
    return np.clip(f, 0.0, 1.0)


def _block(seed, b, grid, nlev, vertical_lev, active, ncol_total, perm, winds=True):
    c0 = b * BLOCK
    n = min(BLOCK, ncol_total - c0)
    g = np.random.Generator(np.random.Philox(key=int(seed), counter=[0, 0, 0, int(b)]))
    ci = np.arange(c0, c0 + n, dtype=np.int64)
    if perm is not None:
        ci = perm[ci]
    un = g.random((6, n))
    ps = 960.0 + 75.0 * _smooth01(ci, grid, 1.0, 2.0, 0.11, un[0])
    if active:
        ts = 8.0 + 28.0 * _smooth01(ci, grid, 1.5, 1.0, 0.23, un[1])
    else:
        ts = -30.0 + 66.0 * _smooth01(ci, grid, 1.5, 1.0, 0.23, un[1])
    dep = 1.0 + 11.0 * _smooth01(ci, grid, 2.0, 3.0, 0.41, un[2])
    tds = np.minimum(ts - dep, 27.0)
    dep = ts - tds
    lapse = 5.5 + 2.5 * _smooth01(ci, grid, 3.0, 1.0, 0.57, un[3])
    ttrop = -60.0 + 8.0 * (2.0 * _smooth01(ci, grid, 1.0, 1.0, 0.71, un[4]) - 1.0)
    dgrad = 2.0 + 4.0 * _smooth01(ci, grid, 2.0, 2.0, 0.83, un[5])

    if vertical_lev == 'pressure':
        p = np.broadcast_to(ERA5_LEVELS_HPA[:nlev].astype(np.float64), (n, nlev))
    else:
        p = ps[:, None] * sigma_levels(nlev)[None, :] * 0.997
    zkm = 44.3308 * (1.0 - (p / ps[:, None]) ** 0.190263)
    # the stream is consumed in order, so drawing only the first two planes gives the same t / td as drawing all four
    nz = g.standard_normal((4 if winds else 2, n, nlev))
    t = np.maximum(ts[:, None] - lapse[:, None] * zkm, ttrop[:, None]) + 0.3 * nz[0]
    td = t - (dep[:, None] + dgrad[:, None] * np.maximum(zkm, 0.0) + np.abs(nz[1]))
    f = np.float32
    out = dict(t=t.astype(f), td=td.astype(f), ps=ps.astype(f), ts=ts.astype(f), tds=tds.astype(f))
    if winds:
        u = 5.0 + 3.0 * zkm + 2.0 * nz[2]
        v = 2.0 + 15.0 * np.sin(zkm / 3.0) + 2.0 * nz[3]
        out.update(u=u.astype(f), v=v.astype(f), us=(0.7 * u[:, 0]).astype(f), vs=(0.7 * v[:, 0]).astype(f))
    if vertical_lev != 'pressure':
        out['p'] = p.astype(f)
    return out


def make_soundings(config='C2', cols=None, active=True, shuffle=False, grid=None, nlev=None,
                   vertical_lev=None, seed=None, winds=True):
    """Synthetic soundings of a named configuration (or explicit grid/nlev/vertical_lev/seed).

    ``cols=(c0, c1)`` restricts to a contiguous block of the flattened grid (a rank's shard).
    Returns a dict of float32 arrays with the vertical axis LAST: ``p`` (``[nlev]`` for pressure
    grids, else ``[ncol, nlev]``), ``t, td, u, v`` ``[ncol, nlev]``, ``ps, ts, tds, us, vs`` ``[ncol]``,
    plus ``vertical_lev``, ``grid``, ``cols``.
    """
    cfg = dict(CONFIGS[config]) if config is not None else {}
    if grid is not None:
        cfg['grid'] = tuple(grid)
    if nlev is not None:
        cfg['nlev'] = int(nlev)
    if vertical_lev is not None:
        cfg['vertical_lev'] = vertical_lev
    if seed is not None:
        cfg['seed'] = int(seed)
    ny, nx = cfg['grid']
    ntot = ny * nx
    c0, c1 = (0, ntot) if cols is None else (int(cols[0]), int(cols[1]))
    if not (0 <= c0 <= c1 <= ntot):
        raise ValueError('cols out of range')
    perm = None
    if shuffle:
        perm = np.random.Generator(np.random.Philox(key=cfg['seed'] + 7919)).permutation(ntot)
    def part(b):
        blk = _block(cfg['seed'], b, cfg['grid'], cfg['nlev'], cfg['vertical_lev'], active, ntot, perm, winds)
        lo, hi = max(c0 - b * BLOCK, 0), min(c1 - b * BLOCK, BLOCK)
        return {k: a[lo:hi] for k, a in blk.items()}
    blocks = range(c0 // BLOCK, (max(c1, c0 + 1) - 1) // BLOCK + 1)
    if len(blocks) >= 8:                      # blocks are independent: numpy releases the GIL in the heavy calls
        import os
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(min(16, len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else 4)) as ex:
            parts = list(ex.map(part, blocks))
    else:
        parts = [part(b) for b in blocks]
    keys = parts[0].keys()
    out = {k: np.concatenate([q[k] for q in parts], axis=0) for k in keys}
    if cfg['vertical_lev'] == 'pressure':
        out['p'] = ERA5_LEVELS_HPA[:cfg['nlev']].copy()
    out.update(vertical_lev=cfg['vertical_lev'], grid=cfg['grid'], cols=(c0, c1))
    return out
