"""The dask (core._calc_*_gufunc, reference core.py:237-258 / 449-469) and xarray (xcape_b200/xarray.py) paths.

dask and xarray are installed neither in the build image nor on the GPU box, so these paths run here against
the stand-ins of tests/stubs.py (same call contracts: one function call per block of the loop dimensions, core
dimension last and never split) — and against the real packages too wherever they can be imported.
CPU tests drive the plumbing with method='dummy' / a fake SRH backend; the `gpu` tests run method='cuda'
block-wise and compare with the plain numpy call.
"""
import importlib
import sys

import numpy as np
import pytest

import stubs
from xcape_b200 import core


def _backends():
    out = ['stub']
    try:
        import dask.array  # noqa: F401
        out.append('real')
    except ImportError:
        pass
    return out


@pytest.fixture(params=_backends())
def da(request, monkeypatch):
    """core.da pointed at the stub (or the real dask.array)."""
    if request.param == 'real':
        import dask.array as real
        monkeypatch.setattr(core, 'da', real)
        return real
    monkeypatch.setattr(core, 'da', stubs.dask_array)
    return stubs.dask_array


@pytest.fixture
def xx(monkeypatch):
    """xcape_b200.xarray imported against the stub xarray (real one if installed), with core.da matching."""
    try:
        import xarray as xr_real  # noqa: F401
        real = True
    except ImportError:
        real = False
    if not real:
        monkeypatch.setitem(sys.modules, 'xarray', stubs.xarray)
        monkeypatch.setattr(core, 'da', stubs.dask_array)
    sys.modules.pop('xcape_b200.xarray', None)
    mod = importlib.import_module('xcape_b200.xarray')
    yield mod
    sys.modules.pop('xcape_b200.xarray', None)


def _fields(ny=6, nx=10, nlev=12, seed=0):
    rng = np.random.default_rng(seed)
    g = (ny, nx)
    p1 = np.linspace(1000, 100, nlev).astype(np.float32)
    d = dict(p=np.broadcast_to(p1, g + (nlev,)).copy(), t=rng.normal(size=g + (nlev,)).astype(np.float32),
             td=rng.normal(size=g + (nlev,)).astype(np.float32), u=rng.normal(size=g + (nlev,)).astype(np.float32),
             v=rng.normal(size=g + (nlev,)).astype(np.float32))
    for k in ('ps', 'ts', 'tds', 'us', 'vs'):
        d[k] = rng.normal(size=g).astype(np.float32)
    d['p1'] = p1
    return d


def _fake_srh_backend(calls):
    """Stands in for srh_cuda.srh_fused on a box without a GPU: outputs encode the inputs so that a block mix-up shows."""
    def srh_fused(p_2d, t_2d, td_2d, u_2d, v_2d, p_s, t_s, td_s, u_s, v_s, flag_1d, plp, depth, aglh0, type_grid, output, **kw):
        n = t_2d.shape[1]
        calls.append((n, flag_1d, type_grid, tuple(np.asarray(p_2d).shape)))
        a = np.asarray(t_2d, np.float64).sum(axis=0) + np.asarray(u_s, np.float64)
        b = np.asarray(td_2d, np.float64).sum(axis=0) - np.asarray(v_s, np.float64)
        if output == 1:
            return a, b
        pair = lambda s: np.stack([np.asarray(u_2d[0], np.float32) * s, np.asarray(v_2d[0], np.float32) * s])
        return a, b, pair(1), pair(2), pair(3)
    return srh_fused


# ------------------------------------------------------------------------------------------------- dask, CPU
@pytest.mark.parametrize('source,n_out', [('surface', 2), ('most-unstable', 4)])
@pytest.mark.parametrize('pressure', [False, True])
def test_cape_gufunc_signature_blocks_and_dtypes(da, source, n_out, pressure):
    d = _fields()
    ch = (2, 5, -1)
    args = [da.from_array(d[k], ch) for k in ('t', 'td')] + [da.from_array(d[k], ch[:2]) for k in ('ps', 'ts', 'tds')]
    p = d['p1'] if pressure else da.from_array(d['p'], ch)
    out = core.calc_cape(p, *args, source=source, method='dummy', vertical_lev='pressure' if pressure else 'sigma')
    assert len(out) == n_out                         # reference core.py:246 lists 7 inputs for 6 arguments here
    for o, dt in zip(out, ('f4', 'f4', 'i4', 'f4')):
        assert isinstance(o, da.Array) and o.shape == (6, 10) and o.dtype == np.dtype(dt)
        assert np.array_equal(np.asarray(o.compute()), np.ones((6, 10)))
    if da is stubs.dask_array:
        assert stubs.apply_gufunc.calls == 3 * 2     # one call per (2, 5) block of the (6, 10) grid


def test_cape_gufunc_rejects_a_split_core_dimension(da):
    d = _fields()
    ch = (2, 5, 6)                                    # level axis in two chunks
    args = [da.from_array(d[k], ch) for k in ('p', 't', 'td')] + [da.from_array(d[k], ch[:2]) for k in ('ps', 'ts', 'tds')]
    with pytest.raises(ValueError):
        r = core.calc_cape(*args, source='surface', method='dummy', vertical_lev='sigma')
        [x.compute() for x in r]


@pytest.mark.parametrize('output_var,n_out', [('srh', 2), ('all', 8)])
@pytest.mark.parametrize('pressure', [False, True])
def test_srh_gufunc_blocks_match_the_unchunked_call(da, monkeypatch, output_var, n_out, pressure):
    from xcape_b200 import srh_cuda
    calls = []
    monkeypatch.setattr(srh_cuda, 'srh_fused', _fake_srh_backend(calls))
    d = _fields()
    ch = (3, 4, -1)
    vl = 'pressure' if pressure else 'sigma'
    names3, names2 = ('t', 'td', 'u', 'v'), ('ps', 'ts', 'tds', 'us', 'vs')
    plain = core.calc_srh(d['p1'] if pressure else d['p'], *(d[k] for k in names3 + names2), vertical_lev=vl, output_var=output_var)
    n_plain = len(calls)
    chunked = core.calc_srh(d['p1'] if pressure else da.from_array(d['p'], ch), *(da.from_array(d[k], ch) for k in names3),
                            *(da.from_array(d[k], ch[:2]) for k in names2), vertical_lev=vl, output_var=output_var)
    assert len(plain) == len(chunked) == n_out
    for a, b, dt in zip(plain, chunked, ('f8', 'f8') + ('f4',) * 6):
        assert isinstance(b, da.Array) and b.dtype == np.dtype(dt)
        assert np.allclose(np.asarray(b.compute()), a)
    blocks = calls[n_plain:]
    if da is stubs.dask_array:
        assert sorted(c[0] for c in blocks) == [6, 6, 12, 12, 12, 12]              # (6, 10) grid in (3, 4) blocks: 2 x 3, last column of blocks ragged
    assert all(c[1] == (1 if pressure else 0) and c[2] == (2 if pressure else 1) for c in calls)
    if pressure:
        assert all(c[3] == (12, 1) for c in calls)                                  # the shared axis reaches every block whole


# ------------------------------------------------------------------------------------------------ xarray, CPU
def test_xarray_layer_names_dims_and_chunked_inputs(xx, monkeypatch):
    import xarray as xr                               # the stub, or the real thing
    from xcape_b200 import srh_cuda
    monkeypatch.setattr(srh_cuda, 'srh_fused', _fake_srh_backend([]))
    d = _fields()
    dims3, dims2 = ('y', 'x', 'level'), ('y', 'x')
    A3 = {k: xr.DataArray(d[k], dims=dims3) for k in ('p', 't', 'td', 'u', 'v')}
    A2 = {k: xr.DataArray(d[k], dims=dims2) for k in ('ps', 'ts', 'tds', 'us', 'vs')}
    ds = xx.calc_cape(A3['p'], A3['t'], A3['td'], A2['ps'], A2['ts'], A2['tds'], source='most-unstable', method='dummy')
    assert list(ds) == ['cape', 'cin', 'mulev', 'zmulev']
    for n in ds:
        assert tuple(ds[n].dims) == dims2 and np.array_equal(np.asarray(ds[n].values), np.ones((6, 10)))
    assert ds['cape'].attrs['units'] == 'J kg-1' and ds.attrs['source'] == 'most-unstable'
    # level axis not last + a 1-D pressure coordinate + chunked inputs
    t_first = xr.DataArray(np.moveaxis(d['t'], -1, 0), dims=('level', 'y', 'x'))
    p1 = xr.DataArray(d['p1'], dims=('level',))
    ds2 = xx.calc_cape(p1, t_first, A3['td'], A2['ps'], A2['ts'], A2['tds'], method='dummy', vertical_lev='pressure')
    assert list(ds2) == ['cape', 'cin'] and tuple(ds2['cin'].dims) == dims2
    ss = xx.calc_srh(*(A3[k] for k in ('p', 't', 'td', 'u', 'v')), *(A2[k] for k in ('ps', 'ts', 'tds', 'us', 'vs')), output_var='all')
    ref = core.calc_srh(*(d[k] for k in ('p', 't', 'td', 'u', 'v', 'ps', 'ts', 'tds', 'us', 'vs')), output_var='all')
    names = ('srh_rm', 'srh_lm', 'rm_u', 'rm_v', 'lm_u', 'lm_v', 'mean_6km_u', 'mean_6km_v')
    assert list(ss) == list(names)
    for n, r in zip(names, ref):
        assert np.allclose(np.asarray(ss[n].values), r)
    ck = {'y': 2, 'x': 5}
    sc = xx.calc_srh(*(A3[k].chunk(ck) for k in ('p', 't', 'td', 'u', 'v')), *(A2[k].chunk(ck) for k in ('ps', 'ts', 'tds', 'us', 'vs')),
                     output_var='srh')
    for n, r in zip(names[:2], ref):
        assert np.allclose(np.asarray(sc[n].values), r)


# ------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize('cfg,vl', [('C1', 'sigma'), ('C2', 'pressure')])
def test_dask_blocks_on_the_gpu_match_the_numpy_call(da, cfg, vl):
    """method='cuda' once per dask block == one call on the whole array, bit for bit (CAPE and SRH)."""
    from xcape_b200.synthetic import make_soundings
    d = make_soundings(cfg, cols=(0, 1000))
    g = (25, 40)
    nlev = d['t'].shape[1]
    f3 = {k: d[k].reshape(g + (nlev,)) for k in ('t', 'td', 'u', 'v')}
    f2 = {k: d[k].reshape(g) for k in ('ps', 'ts', 'tds', 'us', 'vs')}
    p_np = d['p'] if vl == 'pressure' else d['p'].reshape(g + (nlev,))
    ch = (10, 16, -1)
    p_da = p_np if vl == 'pressure' else da.from_array(p_np, ch)
    kw = dict(source='most-unstable', pinc=500., vertical_lev=vl)
    ref = core.calc_cape(p_np, f3['t'], f3['td'], f2['ps'], f2['ts'], f2['tds'], **kw)
    got = core.calc_cape(p_da, da.from_array(f3['t'], ch), da.from_array(f3['td'], ch),
                         *(da.from_array(f2[k], ch[:2]) for k in ('ps', 'ts', 'tds')), **kw)
    for a, b in zip(ref, got):
        assert np.array_equal(a, np.asarray(b.compute()))
    skw = dict(depth=3000, vertical_lev=vl, output_var='all')
    sref = core.calc_srh(p_np, *(f3[k] for k in ('t', 'td', 'u', 'v')), *(f2[k] for k in ('ps', 'ts', 'tds', 'us', 'vs')), **skw)
    sgot = core.calc_srh(p_da, *(da.from_array(f3[k], ch) for k in ('t', 'td', 'u', 'v')),
                         *(da.from_array(f2[k], ch[:2]) for k in ('ps', 'ts', 'tds', 'us', 'vs')), **skw)
    for a, b in zip(sref, sgot):
        assert np.array_equal(np.asarray(a), np.asarray(b.compute()).astype(np.asarray(a).dtype))


@pytest.mark.gpu
def test_xarray_layer_on_the_gpu_matches_the_numpy_call(xx):
    import xarray as xr
    from xcape_b200.synthetic import make_soundings
    d = make_soundings('C1')
    g = (25, 40)
    dims3, dims2 = ('y', 'x', 'level'), ('y', 'x')
    A3 = {k: xr.DataArray(d[k].reshape(g + (50,)), dims=dims3) for k in ('p', 't', 'td', 'u', 'v')}
    A2 = {k: xr.DataArray(d[k].reshape(g), dims=dims2) for k in ('ps', 'ts', 'tds', 'us', 'vs')}
    ref = core.calc_cape(d['p'], d['t'], d['td'], d['ps'], d['ts'], d['tds'], source='most-unstable', vertical_lev='sigma')
    for chunks in (None, {'y': 5, 'x': 8}):
        c = (lambda a: a) if chunks is None else (lambda a: a.chunk(chunks))
        ds = xx.calc_cape(c(A3['p']), c(A3['t']), c(A3['td']), c(A2['ps']), c(A2['ts']), c(A2['tds']), source='most-unstable',
                          vertical_lev='sigma')
        for name, r in zip(('cape', 'cin', 'mulev', 'zmulev'), ref):
            assert tuple(ds[name].dims) == dims2 and np.array_equal(np.asarray(ds[name].values).ravel(), r)
    ss = xx.calc_srh(*(A3[k] for k in ('p', 't', 'td', 'u', 'v')), *(A2[k] for k in ('ps', 'ts', 'tds', 'us', 'vs')),
                     output_var='all', vertical_lev='sigma')
    sref = core.calc_srh(*(d[k] for k in ('p', 't', 'td', 'u', 'v', 'ps', 'ts', 'tds', 'us', 'vs')), output_var='all',
                         vertical_lev='sigma')
    for name, r in zip(('srh_rm', 'srh_lm', 'rm_u', 'rm_v', 'lm_u', 'lm_v', 'mean_6km_u', 'mean_6km_v'), sref):
        assert np.array_equal(np.asarray(ss[name].values).ravel(), np.asarray(r).ravel())
