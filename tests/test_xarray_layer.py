"""xarray front end (optional dependency; skipped where xarray is not installed — the build image
does not have it).  Needs a GPU because it calls method='cuda'."""
import numpy as np
import pytest

xr = pytest.importorskip('xarray')
pytestmark = pytest.mark.gpu


def test_xarray_cape_and_srh_match_numpy_api():
    from xcape_b200 import core
    from xcape_b200 import xarray as xx
    from xcape_b200.synthetic import make_soundings
    d = make_soundings('C1')
    g = (25, 40)
    dims3, dims2 = ('y', 'x', 'level'), ('y', 'x')
    da3 = {k: xr.DataArray(d[k].reshape(g + (50,)), dims=dims3) for k in ('p', 't', 'td', 'u', 'v')}
    da2 = {k: xr.DataArray(d[k].reshape(g), dims=dims2) for k in ('ps', 'ts', 'tds', 'us', 'vs')}
    ds = xx.calc_cape(da3['p'], da3['t'], da3['td'], da2['ps'], da2['ts'], da2['tds'], source='most-unstable',
                      vertical_lev='sigma')
    ref = core.calc_cape(d['p'], d['t'], d['td'], d['ps'], d['ts'], d['tds'], source='most-unstable', vertical_lev='sigma')
    for name, r in zip(('cape', 'cin', 'mulev', 'zmulev'), ref):
        assert ds[name].dims == dims2 and np.array_equal(ds[name].values.ravel(), r)
    ss = xx.calc_srh(*(da3[k] for k in ('p', 't', 'td', 'u', 'v')), *(da2[k] for k in ('ps', 'ts', 'tds', 'us', 'vs')),
                     output_var='all', vertical_lev='sigma')
    sref = core.calc_srh(*(d[k] for k in ('p', 't', 'td', 'u', 'v', 'ps', 'ts', 'tds', 'us', 'vs')), output_var='all',
                         vertical_lev='sigma')
    for name, r in zip(('srh_rm', 'srh_lm', 'rm_u', 'rm_v', 'lm_u', 'lm_v', 'mean_6km_u', 'mean_6km_v'), sref):
        assert np.array_equal(ss[name].values.ravel(), r)
