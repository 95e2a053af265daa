"""CUDA counterpart of the reference's ``stdheight.py`` shim (stdheight.py:5-38): hypsometric
height above ground of every level, float64, via ``xcape_cuda_stdheight``."""
import ctypes as C

import numpy as np

from . import _array as A
from . import _lib


def stdheight(p_2d, t_2d, td_2d, p_s, t_s, td_s, flag_1d, pres_lev_pos, aglh0, type_grid, *,
              device=0, stream=None, top_first=False):
    """Same arguments as ``xcape.stdheight.stdheight``; returns ``(H2D (nlev, ngrid) float64,
    H_s (ngrid,) float64)``.  Levels below ``pres_lev_pos`` are -999999
    (stdheight_2D_pressure_lev.f90:85-87).  ``aglh0`` must be a scalar here.  ``top_first``:
    the level axis of the inputs is stored model top first; ``H2D`` comes back in that same order
    (``pres_lev_pos`` counts from the surface)."""
    L = _lib.lib()
    nlev, ngrid = t_2d.shape
    if not np.isscalar(aglh0):
        raise ValueError('aglh0 must be a scalar height (m)')
    if type_grid == 1:
        p_is_1d = 0
        f3, f1, _, dt, layout, mem, ref = A.prepare_fields([p_2d, t_2d, td_2d], [p_s, t_s, td_s])
        p, t_, td_ = f3
    elif type_grid == 2 and flag_1d == 1:
        p_is_1d = 1
        f3, f1, p, dt, layout, mem, ref = A.prepare_fields([t_2d, td_2d], [p_s, t_s, td_s], p=p_2d)
        t_, td_ = f3
    else:
        raise ValueError('type_grid must be 1, or 2 with a 1-D pressure array')
    ps_, ts_, tds_ = f1
    start = None
    if p_is_1d and pres_lev_pos is not None:
        if A.is_cuda(ref):
            import torch
            start = torch.as_tensor(pres_lev_pos, device=ref.device).to(torch.int32).expand(ngrid).contiguous()
        else:
            start = np.ascontiguousarray(np.broadcast_to(np.asarray(pres_lev_pos), (ngrid,)), dtype=np.int32)
    if layout == _lib.LEVEL_MAJOR:
        h = A.empty_like_host_or_device(ref, (nlev, ngrid), 'float64')
        h_view = h
    else:
        h = A.empty_like_host_or_device(ref, (ngrid, nlev), 'float64')
        h_view = h.t() if A.is_cuda(ref) else h.T
    hs = A.empty_like_host_or_device(ref, (ngrid,), 'float64')
    A.order_on_stream(ref, stream, [p, t_, td_, ps_, ts_, tds_, start, h, hs])
    rc = L.xcape_cuda_stdheight(A.ptr(p), A.ptr(t_), A.ptr(td_), A.ptr(ps_), A.ptr(ts_), A.ptr(tds_),
                                C.c_int64(ngrid), nlev, p_is_1d, dt, layout | (_lib.LEVELS_TOP_FIRST if top_first else 0), mem,
                                C.c_double(float(aglh0)),
                                A.ptr(start), A.ptr(h), A.ptr(hs), A.device_of(ref, device),
                                A.stream_of(ref, stream))
    _lib.check(rc)
    return h_view, hs
