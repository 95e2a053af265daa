"""numpy / dask / torch front end: ``calc_cape`` and ``calc_srh`` with ``method='cuda'``.

This module mirrors the public surface of the reference's ``xcape/core.py`` (calc_cape
core.py:130-235, calc_srh core.py:335-447 and their ``_calc_*_gufunc`` / ``_calc_*_numpy``
helpers) so that it can stand in for it: same positional arguments, same keyword names and
meanings, same number / shape / dtype of results, same ``ValueError`` / ``KeyError`` behaviour for
bad arguments.  What changes is what runs underneath:

=====================  ==========================================  ==============================
step                   reference                                   here (``method='cuda'``)
=====================  ==========================================  ==============================
flatten to columns     transposed views, (nlev, ncol)              same views, passed zero-copy
``pres_lev_pos``       numpy masked argmin, (nlev, ncol) float64   on the device, in-kernel glue
                       temporary (core.py:286-289)
dtype coercion         f2py copies (float64 -> float32 ...)        cast fused into the relayout
CAPE / SRH columns     serial Fortran loop per dask block          sm_100a kernels, thread/column
=====================  ==========================================  ==============================

Differences that are deliberate (and documented in DESIGN.md): ``method`` defaults to
``'cuda'`` (the Fortran extension cannot be built in this environment; ``method='fortran'``
forwards to an installed reference ``xcape`` if there is one); ``calc_srh`` gains ``method=``;
``vertical_lev`` may be omitted (defaults to ``'sigma'`` as the reference's docstring promises but
its code does not, core.py:229); inputs may be torch CUDA tensors (results are then CUDA tensors);
``lev_axis=0`` accepts level-major ``[nlev, ...]`` arrays (the on-disk order of ERA5 / HRRR) with
zero relayout; ``device=`` / ``devices=[...]`` pick the GPU(s); ``calc_cape(precision='fast')`` trades
bit-exactness of CAPE/CIN for about four times the speed (tolerance-level parity, MU level still exact;
``'fast-relaxed'`` keeps the reference's iteration and is about 1.8x faster).
"""
from functools import reduce

import numpy as np

try:  # dask is optional here (the reference imports it unconditionally, core.py:9)
    import dask.array as da
except ImportError:  # pragma: no cover - exercised only where dask is absent
    da = None

from . import _array as A

_SOURCE = {'surface': 1, 'most-unstable': 2, 'mixed-layer': 3}            # core.py:302
_ADIABAT = {'pseudo-liquid': 1, 'reversible-liquid': 2,                    # core.py:303-304
            'pseudo-ice': 3, 'reversible-ice': 4}
_VERTICAL = {'sigma': 1, 'pressure': 2}                                    # core.py:305
_OUTPUT = {'srh': 1, 'all': 2}                                             # core.py:511


def _prod(v):
    return reduce(lambda x, y: x * y, v, 1)


def _shape(a):
    return tuple(a.shape)


def _as_array(a):
    return a if A.is_torch(a) else np.asarray(a)


def _columns_2d(arrays, lev_axis):
    """``[..., nlev]`` (or ``[nlev, ...]``) -> ``(nlev, ncol)`` views.  Same contract as the
    reference's ``_reshape_inputs`` (core.py:31-50): all shapes equal, no data copied for
    dense inputs."""
    shp = _shape(arrays[0])
    for a in arrays:
        if _shape(a) != shp:
            raise ValueError('Input arrays must have the same shape.')
    out = []
    for a in arrays:
        if a.ndim == 1:
            a = a.reshape(1, -1) if lev_axis != 0 else a.reshape(-1, 1)
        if lev_axis == 0:
            out.append(a.reshape(a.shape[0], -1))
        else:
            a2 = a.reshape(-1, a.shape[-1])
            out.append(a2.t() if A.is_torch(a2) else a2.transpose())
    return out


def _columns_1d(arrays):
    """Surface fields -> ``(ncol,)`` (reference ``_reshape_surface_inputs``, core.py:81-100)."""
    shp = _shape(arrays[0])
    for a in arrays:
        if _shape(a) != shp:
            raise ValueError('Input arrays must have the same shape.')
    return [a.reshape(-1) for a in arrays]


def _grid_shape(field_shape, lev_axis):
    if len(field_shape) == 1:
        return (1,)
    return field_shape[1:] if lev_axis == 0 else field_shape[:-1]


def _unflatten(arrays, grid_shape):
    return [a.reshape(grid_shape) for a in arrays]


_LEVEL_ORDERS = ('surface_first', 'top_first', 'auto')


def _top_first(p, lev_axis, level_order):
    """Is the level axis stored model top first?  The reference assumes surface -> top and leaves the
    flip of e.g. ERA5 downloads (1 -> 1000 hPa) to the caller; here
    ``level_order='top_first'`` declares it and ``'auto'`` reads it off the pressure array (first
    column of an N-D ``p``): pressure increasing along the axis means top first."""
    if level_order not in _LEVEL_ORDERS:
        raise ValueError(f"`level_order` must be one of: {list(_LEVEL_ORDERS)}")
    if level_order != 'auto':
        return level_order == 'top_first'
    if A.is_torch(p):
        p = p.detach()
    if p.ndim == 1:
        first, last = float(p[0]), float(p[-1])
    else:
        # first / last level of a column whose two values are both finite (masked points carry NaN or fill values):
        # look at a bounded sample of columns rather than at column 0 only
        col = p.reshape(p.shape[0], -1) if lev_axis == 0 else p.reshape(-1, p.shape[-1]).T
        ncol = col.shape[1]
        idx = np.unique(np.linspace(0, ncol - 1, num=min(ncol, 257)).astype(np.int64))
        lo, hi = col[0][idx], col[-1][idx]
        lo, hi = (np.asarray(x.cpu() if A.is_torch(x) else x, dtype=np.float64) for x in (lo, hi))
        ok = np.isfinite(lo) & np.isfinite(hi) & (np.abs(lo) < 1e6) & (np.abs(hi) < 1e6) & (lo != hi)
        if not ok.any():
            raise ValueError("level_order='auto': no column with finite, distinct first and last pressures in the sample; "
                             "pass level_order='surface_first' or 'top_first'")
        first, last = float(lo[ok][0]), float(hi[ok][0])
    if not (np.isfinite(first) and np.isfinite(last)) or first == last:
        raise ValueError("level_order='auto': the pressure axis does not determine the order; pass level_order explicitly")
    return bool(first < last)


def _any_dask_array(*args):
    return da is not None and any(isinstance(a, da.Array) for a in args)


def _concrete(a):
    """numpy array of a small argument that may be chunked (the shared 1-D pressure axis)."""
    return np.asarray(a.compute() if hasattr(a, 'compute') else a)


def _cape_dummy(*args, **kwargs):
    """The reference's fake backend for shape tests (core.py:107-121)."""
    p, t, td, ps, ts, tds = args
    assert p.ndim == 2
    n = t.shape[1]
    return tuple(np.ones((1, n)) for _ in range(4))


def _reference_shims():
    try:
        from xcape.cape_fortran import cape as cape_f
        from xcape.srh import srh as srh_f
        from xcape.stdheight import stdheight as stdh_f
    except Exception as e:  # noqa: BLE001
        raise ImportError("method='fortran' needs the reference xcape package with its compiled f2py "
                          "extensions on sys.path; it is not part of xcape_b200") from e
    return cape_f, srh_f, stdh_f


# --------------------------------------------------------------------------------------
# CAPE
# --------------------------------------------------------------------------------------
def calc_cape(*args, **kwargs):
    """Convective available potential energy and convective inhibition of every column.

    ``calc_cape(p, t, td, ps, ts, tds, source='surface', ml_depth=500., adiabat='pseudo-liquid',
    pinc=500., method='cuda', vertical_lev='sigma')`` — see the reference docstring
    (core.py:131-219) for the physics.  ``p, t, td``: hPa / degC / degC with the vertical axis
    LAST (``p`` is 1-D ``[nlev]`` for ``vertical_lev='pressure'``); ``ps, ts, tds``: surface values,
    shape ``t.shape[:-1]``.

    Returns ``(cape, cin)`` or, for ``source='most-unstable'``, ``(cape, cin, MUlev, zMUlev)``;
    float32 except ``MUlev`` (int32); shape ``t.shape[:-1]``.
    """
    if len(args) < 6:
        raise ValueError("Too few arguments.")
    if len(args) > 6:
        raise ValueError("Too many arguments.")
    allowed = list(_VERTICAL)
    if kwargs.get('vertical_lev', 'sigma') not in allowed:
        raise ValueError(f"`vertical_lev` must be one of: {allowed}")
    if _any_dask_array(*args):
        return _calc_cape_gufunc(*args, **kwargs)
    return _calc_cape_numpy(*args, **kwargs)


def _calc_cape_gufunc(*args, **kwargs):
    """dask path: one ``_calc_cape_numpy`` call per block (core.py:237-258).  The signature is
    built from the actual arity (the reference lists 7 inputs for 6 arguments on pressure
    grids, core.py:246)."""
    p_is_1d = (args[0].ndim == 1)
    sig_in = ['(i)'] * 3 + ['()'] * 3
    n_out = 4 if kwargs.get('source', 'surface') == 'most-unstable' else 2
    signature = ','.join(sig_in) + '->' + ','.join(['()'] * n_out)
    dtypes = ('f4', 'f4', 'i4', 'f4')[:n_out]
    if p_is_1d:
        # a 1-D pressure axis is shared by every block: bind it instead of broadcasting it
        p = _concrete(args[0])
        sig = ','.join(['(i)'] * 2 + ['()'] * 3) + '->' + ','.join(['()'] * n_out)
        return da.apply_gufunc(lambda t, td, ps, ts, tds, **kw: _calc_cape_numpy(p, t, td, ps, ts, tds, **kw),
                               sig, *args[1:], output_dtypes=dtypes, axis=-1, vectorize=False, **kwargs)
    return da.apply_gufunc(_calc_cape_numpy, signature, *args, output_dtypes=dtypes, axis=-1,
                           vectorize=False, **kwargs)


def _calc_cape_numpy(*args, source='surface', ml_depth=500., adiabat='pseudo-liquid', pinc=500.,
                     method='cuda', vertical_lev='sigma', lev_axis=-1, device=0, devices=None,
                     stream=None, precision='faithful', level_order='surface_first'):
    """Flatten to columns, dispatch on ``method``, restore the grid shape (core.py:261-332)."""
    p, t, td, ps, ts, tds = (_as_array(a) for a in args)
    lev_axis = 0 if lev_axis == 0 else -1
    top_first = _top_first(p, lev_axis, level_order)
    grid_shape = _grid_shape(_shape(t), lev_axis)

    p_s1d, t_s1d, td_s1d = _columns_1d([ps, ts, tds])
    if p.ndim == 1 and (t.ndim > 1 or vertical_lev == 'pressure'):
        t_2d, td_2d = _columns_2d([t, td], lev_axis)
        p_2d = p.reshape(-1, 1)
        flag_1d = 1
    elif _shape(p) == _shape(t) and vertical_lev == 'sigma':
        p_2d, t_2d, td_2d = _columns_2d([p, t, td], lev_axis)
        flag_1d = 0
    elif _shape(p) == _shape(t) and vertical_lev == 'pressure':
        raise ValueError("P should be 1d")
    else:
        raise ValueError('Input arrays must have the same shape.')

    opts = dict(source=_SOURCE[source], ml_depth=ml_depth, adiabat=_ADIABAT[adiabat], pinc=pinc,
                type_grid=_VERTICAL[vertical_lev])

    if method == 'cuda':
        from .cape_cuda import cape as _cape_cuda
        # pres_lev_pos=None: computed on the device instead of core.py:286-289's numpy temporaries
        outs = _cape_cuda(p_2d, t_2d, td_2d, p_s1d, t_s1d, td_s1d, flag_1d, None, **opts,
                          device=device, devices=devices, stream=stream, precision=precision,
                          top_first=top_first)
    elif method in ('fortran', 'dummy'):
        host = [A.to_host_numpy(a) for a in (p_2d, t_2d, td_2d, p_s1d, t_s1d, td_s1d)]
        if top_first:
            host[:3] = [np.ascontiguousarray(a[::-1]) for a in host[:3]]
        if method == 'dummy':
            outs = _cape_dummy(*host, **opts)
        else:
            cape_f, _, _ = _reference_shims()
            plp = 1
            if flag_1d:
                plp = np.ma.masked_less(host[3] - host[0], 0).argmin(axis=0) + 1
            outs = cape_f(*host, flag_1d, plp, **opts)
    else:
        raise ValueError('invalid method')

    n_out = 4 if _SOURCE[source] == 2 else 2
    return tuple(_unflatten(list(outs[:n_out]), grid_shape))


# --------------------------------------------------------------------------------------
# SRH
# --------------------------------------------------------------------------------------
def calc_srh(*args, **kwargs):
    """Storm-relative helicity (right- and left-moving Bunkers storms) of every column.

    ``calc_srh(p, t, td, u, v, ps, ts, tds, us, vs, depth=3000, vertical_lev='sigma',
    output_var='srh', method='cuda')`` — physics as in the reference docstring (core.py:336-443).
    Returns ``(srh_rm, srh_lm)`` or, for ``output_var='all'``, additionally ``rm_u, rm_v, lm_u,
    lm_v, mean_6km_u, mean_6km_v``; shape ``t.shape[:-1]``; srh float64, the rest float32.
    """
    if len(args) != 10:
        raise ValueError("calc_srh takes 10 positional arrays: p, t, td, u, v, ps, ts, tds, us, vs")
    if kwargs.get('vertical_lev', 'sigma') not in _VERTICAL:
        raise ValueError(f"`vertical_lev` must be one of: {list(_VERTICAL)}")
    if _any_dask_array(*args):
        return _calc_srh_gufunc(*args, **kwargs)
    return _calc_srh_numpy(*args, **kwargs)


def _calc_srh_gufunc(*args, **kwargs):
    """dask path (core.py:449-469), arity-correct signature."""
    n_out = 8 if kwargs.get('output_var', 'srh') == 'all' else 2
    dtypes = ('f8', 'f8') + ('f4',) * 6
    outs = ','.join(['()'] * n_out)
    if args[0].ndim == 1:
        p = _concrete(args[0])
        sig = ','.join(['(i)'] * 4 + ['()'] * 5) + '->' + outs
        return da.apply_gufunc(lambda t, td, u, v, ps, ts, tds, us, vs, **kw:
                               _calc_srh_numpy(p, t, td, u, v, ps, ts, tds, us, vs, **kw),
                               sig, *args[1:], output_dtypes=dtypes[:n_out], axis=-1, vectorize=False, **kwargs)
    sig = ','.join(['(i)'] * 5 + ['()'] * 5) + '->' + outs
    return da.apply_gufunc(_calc_srh_numpy, sig, *args, output_dtypes=dtypes[:n_out], axis=-1,
                           vectorize=False, **kwargs)


def _calc_srh_numpy(*args, depth=3000, vertical_lev='sigma', output_var='srh', method='cuda',
                    lev_axis=-1, device=0, devices=None, stream=None, precision='faithful',
                    level_order='surface_first'):
    """Flatten, dispatch, unflatten (core.py:473-542).  ``aglh0 = 2.`` as in core.py:519."""
    p, t, td, u, v, ps, ts, tds, us, vs = (_as_array(a) for a in args)
    lev_axis = 0 if lev_axis == 0 else -1
    top_first = _top_first(p, lev_axis, level_order)
    grid_shape = _grid_shape(_shape(t), lev_axis)
    surf = _columns_1d([ps, ts, tds, us, vs])
    if p.ndim == 1 and (t.ndim > 1 or vertical_lev == 'pressure'):
        t_2d, td_2d, u_2d, v_2d = _columns_2d([t, td, u, v], lev_axis)
        p_2d = p.reshape(-1, 1)
        flag_1d = 1
    elif _shape(p) == _shape(t) and vertical_lev == 'sigma':
        p_2d, t_2d, td_2d, u_2d, v_2d = _columns_2d([p, t, td, u, v], lev_axis)
        flag_1d = 0
    elif _shape(p) == _shape(t) and vertical_lev == 'pressure':
        raise ValueError("P should be 1d")
    else:
        raise ValueError('Input arrays must have the same shape.')
    type_grid = _VERTICAL[vertical_lev]
    output = _OUTPUT[output_var]

    if method == 'cuda':
        from .srh_cuda import srh_fused
        outs = srh_fused(p_2d, t_2d, td_2d, u_2d, v_2d, *surf, flag_1d, None, depth, 2., type_grid, output,
                         device=device, devices=devices, stream=stream, precision=precision,
                         top_first=top_first)
    elif method == 'fortran':
        _, srh_f, stdh_f = _reference_shims()
        host = [A.to_host_numpy(a) for a in (p_2d, t_2d, td_2d, u_2d, v_2d, *surf)]
        if top_first:
            host[:5] = [np.ascontiguousarray(a[::-1]) for a in host[:5]]
        plp = 1
        if flag_1d:
            plp = np.ma.masked_less(host[5] - host[0], 0).argmin(axis=0) + 1
        aglh_2d, aglh_s = stdh_f(host[0], host[1], host[2], host[5], host[6], host[7], flag_1d, plp,
                                 aglh0=2., type_grid=type_grid)
        outs = srh_f(host[3], host[4], aglh_2d, host[8], host[9], aglh_s, plp, depth,
                     type_grid=type_grid, output=output)
    else:
        raise ValueError('invalid method')

    srh_rm, srh_lm = _unflatten(list(outs[:2]), grid_shape)
    if output == 1:
        return srh_rm, srh_lm
    comps = []
    for uv in outs[2:5]:                      # (2, ncol): [0] = u component, [1] = v (core.py:541-542)
        comps += _unflatten([uv[0], uv[1]], grid_shape)
    return (srh_rm, srh_lm, *comps)
