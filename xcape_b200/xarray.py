"""xarray front end.

The reference ships ``xcape/xarray.py`` as an un-importable stub (a stale copy of xhistogram's
``histogram``; ``xarray.py:8`` imports a name ``core`` does not define) while its package exports it
(``__init__.py:5``) and its docs reserve a section for it (``doc/api.rst:19-23``).  This module is the
layer that stub points at: ``xr.apply_ufunc`` wrappers of :func:`xcape_b200.core.calc_cape` and
:func:`xcape_b200.core.calc_srh` with the vertical dimension as the core dimension, named outputs,
and ``dask='parallelized'`` so that chunked datasets map one block to one kernel call.

xarray is an optional dependency: importing this module without it raises ``ImportError``.
"""
import numpy as np

try:
    import xarray as xr
except ImportError as e:  # pragma: no cover - xarray is not installed in the build image
    raise ImportError('xcape_b200.xarray needs the optional dependency xarray') from e

from . import core as _core

__all__ = ['calc_cape', 'calc_srh']


def _lev_dims(arr, lev_dim):
    return [lev_dim] if lev_dim in arr.dims else []


def calc_cape(p, t, td, ps, ts, tds, *, lev_dim='level', source='surface', ml_depth=500.,
              adiabat='pseudo-liquid', pinc=500., method='cuda', vertical_lev='sigma', **kwargs):
    """CAPE / CIN of every column of ``xarray.DataArray`` inputs.

    ``p, t, td`` carry ``lev_dim`` (``p`` may be 1-D ``[lev_dim]`` for ``vertical_lev='pressure'``);
    ``ps, ts, tds`` are the surface fields on the remaining dimensions.  Units and keywords as in
    :func:`xcape_b200.core.calc_cape`.  Returns an ``xarray.Dataset`` with ``cape``, ``cin`` (J/kg) and,
    for ``source='most-unstable'``, ``mulev`` (int32) and ``zmulev`` (m).
    """
    names = ['cape', 'cin'] + (['mulev', 'zmulev'] if source == 'most-unstable' else [])
    dtypes = [np.float32, np.float32, np.int32, np.float32][:len(names)]
    kw = dict(source=source, ml_depth=ml_depth, adiabat=adiabat, pinc=pinc, method=method,
              vertical_lev=vertical_lev, **kwargs)
    outs = xr.apply_ufunc(
        _core._calc_cape_numpy, p, t, td, ps, ts, tds, kwargs=kw,
        input_core_dims=[_lev_dims(p, lev_dim), [lev_dim], [lev_dim], [], [], []],
        output_core_dims=[[] for _ in names], dask='parallelized', output_dtypes=dtypes,
        dask_gufunc_kwargs=dict(allow_rechunk=False))
    ds = xr.Dataset({n: o for n, o in zip(names, outs)})
    ds['cape'].attrs.update(units='J kg-1', long_name='convective available potential energy')
    ds['cin'].attrs.update(units='J kg-1', long_name='convective inhibition')
    if 'mulev' in ds:
        ds['mulev'].attrs.update(long_name='most-unstable level (1 = surface, 2 = first used level)')
        ds['zmulev'].attrs.update(units='m', long_name='height above ground of the last level the ascent reached')
    ds.attrs.update(source=source, adiabat=adiabat, pinc=float(pinc), vertical_lev=vertical_lev)
    return ds


def calc_srh(p, t, td, u, v, ps, ts, tds, us, vs, *, lev_dim='level', depth=3000, vertical_lev='sigma',
             output_var='srh', method='cuda', **kwargs):
    """Storm-relative helicity (Bunkers right / left movers) of every column of ``DataArray`` inputs.

    Returns an ``xarray.Dataset`` with ``srh_rm``, ``srh_lm`` (m2/s2) and, for ``output_var='all'``,
    ``rm_u, rm_v, lm_u, lm_v, mean_6km_u, mean_6km_v`` (m/s).
    """
    names = ['srh_rm', 'srh_lm']
    if output_var == 'all':
        names += ['rm_u', 'rm_v', 'lm_u', 'lm_v', 'mean_6km_u', 'mean_6km_v']
    dtypes = [np.float64, np.float64] + [np.float32] * (len(names) - 2)
    kw = dict(depth=depth, vertical_lev=vertical_lev, output_var=output_var, method=method, **kwargs)
    outs = xr.apply_ufunc(
        _core._calc_srh_numpy, p, t, td, u, v, ps, ts, tds, us, vs, kwargs=kw,
        input_core_dims=[_lev_dims(p, lev_dim)] + [[lev_dim]] * 4 + [[]] * 5,
        output_core_dims=[[] for _ in names], dask='parallelized', output_dtypes=dtypes,
        dask_gufunc_kwargs=dict(allow_rechunk=False))
    ds = xr.Dataset({n: o for n, o in zip(names, outs)})
    for n in names[:2]:
        ds[n].attrs.update(units='m2 s-2', long_name=f'0-{depth} m storm-relative helicity')
    for n in names[2:]:
        ds[n].attrs.update(units='m s-1')
    ds.attrs.update(depth=float(depth), vertical_lev=vertical_lev)
    return ds
