"""CUDA counterpart of the reference's ``srh.py`` + ``stdheight.py`` shims.

The reference computes SRH with three f2py calls that exchange two full float64 height
arrays (core.py:516-535 -> stdheight.py:20-36 -> srh.py:41-61).  ``srh_fused`` makes ONE call
to ``xcape_cuda_srh`` (include/xcape_b200.h), which produces heights, Bunkers storm motion
and both helicity integrals in a single pass over the column.  ``srh`` keeps the reference's
two-call form (same signature as ``xcape.srh.srh``, heights supplied by the caller, e.g. from
``stdheight_cuda.stdheight``) through ``xcape_cuda_srh_from_heights``.
"""
import ctypes as C

import numpy as np

from . import _array as A
from . import _lib


def srh_fused(p_2d, t_2d, td_2d, u_2d, v_2d, p_s, t_s, td_s, u_s, v_s, flag_1d, pres_lev_pos, depth,
              aglh0, type_grid, output, *, device=0, devices=None, stream=None, precision='faithful',
              top_first=False):
    """
    Arguments are the union of ``stdheight.stdheight`` (stdheight.py:5) and ``srh.srh``
    (srh.py:4): ``*_2d`` are ``(nlev, ngrid)``, ``*_s`` are ``(ngrid,)``; ``pres_lev_pos`` may be
    ``None`` (computed on the device for pressure grids); ``aglh0`` is the scalar height of the
    surface level (core.py:519 passes 2.0); ``output`` 1 -> ``(srh_rm, srh_lm)``, otherwise
    ``(srh_rm, srh_lm, rm(2, ngrid), lm(2, ngrid), mean_6km(2, ngrid))`` like srh.py:63-66.
    srh_* are float64, the storm-motion arrays float32.  ``precision='fast'``
    evaluates the hypsometric height chain in binary32 (within ~1e-3 m2/s2 of the reference).
    ``top_first``: the level axis is stored model top first and walked backwards on the device.
    """
    L = _lib.lib()
    if precision not in ('faithful', 'fast'):
        raise ValueError("precision must be 'faithful' or 'fast'")
    prec = _lib.PRECISION[precision]
    nlev, ngrid = t_2d.shape
    if type_grid == 1:
        p_is_1d = 0
    elif type_grid == 2:
        if flag_1d != 1:
            raise ValueError('pressure-level grids need a 1-D pressure array (flag_1d == 1)')
        p_is_1d = 1
    else:
        raise ValueError('type_grid must be 1 (model levels) or 2 (pressure levels)')
    if not np.isscalar(aglh0):
        raise ValueError('aglh0 must be a scalar height (m)')

    surf = [p_s, t_s, td_s, u_s, v_s]
    if p_is_1d:
        f3, f1, p, dt, layout, mem, ref = A.prepare_fields([t_2d, td_2d, u_2d, v_2d], surf, p=p_2d)
        t_, td_, u_, v_ = f3
        if p.shape[0] != nlev:
            raise ValueError('p must have nlev entries')
    else:
        f3, f1, _, dt, layout, mem, ref = A.prepare_fields([p_2d, t_2d, td_2d, u_2d, v_2d], surf)
        p, t_, td_, u_, v_ = f3
    if any(tuple(a.shape) != (nlev, ngrid) for a in f3) or any(a.shape[0] != ngrid for a in f1):
        raise ValueError('Input arrays must have the same shape.')
    ps_, ts_, tds_, us_, vs_ = f1

    start = None
    if p_is_1d and pres_lev_pos is not None:
        if A.is_cuda(ref):
            import torch
            start = torch.as_tensor(pres_lev_pos, device=ref.device).to(torch.int32).expand(ngrid).contiguous()
        else:
            start = np.ascontiguousarray(np.broadcast_to(np.asarray(pres_lev_pos), (ngrid,)), dtype=np.int32)

    want_all = (output != 1)
    srm = A.empty_like_host_or_device(ref, (ngrid,), 'float64')
    slm = A.empty_like_host_or_device(ref, (ngrid,), 'float64')
    # (2, ngrid) in Fortran order == (ngrid, 2) C order
    rm = A.empty_like_host_or_device(ref, (ngrid, 2), 'float32') if want_all else None
    lm = A.empty_like_host_or_device(ref, (ngrid, 2), 'float32') if want_all else None
    m6 = A.empty_like_host_or_device(ref, (ngrid, 2), 'float32') if want_all else None
    es = 4 if dt == _lib.F32 else 8
    A.order_on_stream(ref, stream, [p, t_, td_, u_, v_, ps_, ts_, tds_, us_, vs_, start, srm, slm, rm, lm, m6])

    def call(c0, c1, dev):
        n = c1 - c0
        if n <= 0:
            return
        whole = (c0 == 0 and c1 == ngrid)
        keep = []

        def off3(a):
            if layout == _lib.LEVEL_LAST:
                return A.ptr(a) + c0 * nlev * es
            if whole:
                return A.ptr(a)
            keep.append(np.ascontiguousarray(a[:, c0:c1]))     # dense [nlev, n] shard of a level-major host array
            return A.ptr(keep[-1])

        def off1(a, e):
            return None if a is None else A.ptr(a) + c0 * e

        rc = L.xcape_cuda_srh(
            A.ptr(p) if p_is_1d else off3(p), off3(t_), off3(td_), off3(u_), off3(v_),
            off1(ps_, es), off1(ts_, es), off1(tds_, es), off1(us_, es), off1(vs_, es),
            C.c_int64(n), nlev, p_is_1d, dt, layout | (_lib.LEVELS_TOP_FIRST if top_first else 0), mem,
            C.c_double(float(depth)), C.c_double(float(aglh0)),
            off1(start, 4), off1(srm, 8), off1(slm, 8), off1(rm, 8), off1(lm, 8), off1(m6, 8),
            prec, dev, A.stream_of(ref, stream))
        _lib.check(rc)

    if mem == _lib.MEM_HOST and devices is not None and len(devices) > 1:
        # one C call: contiguous 128-aligned column blocks, one host thread per device (xcape_cuda_srh_multi)
        opt = lambda a: None if a is None else A.ptr(a)
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        rc = L.xcape_cuda_srh_multi(
            A.ptr(p), A.ptr(t_), A.ptr(td_), A.ptr(u_), A.ptr(v_), A.ptr(ps_), A.ptr(ts_), A.ptr(tds_), A.ptr(us_), A.ptr(vs_),
            C.c_int64(ngrid), nlev, p_is_1d, dt, layout | (_lib.LEVELS_TOP_FIRST if top_first else 0),
            C.c_double(float(depth)), C.c_double(float(aglh0)), opt(start), A.ptr(srm), A.ptr(slm), opt(rm), opt(lm), opt(m6),
            prec, devs, len(devices))
        _lib.check(rc)
    else:
        call(0, ngrid, A.device_of(ref, devices[0] if devices else device))

    if not want_all:
        return srm, slm
    # hand back the reference's (2, ngrid) view (srh.py:42-43: rm_sup[0, :] is the u component)
    tr = (lambda a: a.t()) if A.is_cuda(ref) else (lambda a: a.T)
    return srm, slm, tr(rm), tr(lm), tr(m6)


def srh(u_2d, v_2d, aglh_2d, u_s, v_s, aglh_s, pres_lev_pos, depth, type_grid, output, *, device=0, stream=None,
        top_first=False):
    """Same arguments and returns as ``xcape.srh.srh`` (srh.py:4-66): winds and heights ``(nlev,
    ngrid)``, surface values ``(ngrid,)``; ``pres_lev_pos`` (1-based first level used) only matters
    for ``type_grid == 2``.  Replaces ``bunkers_loop_*`` + ``loop_sreh_*``.  ``top_first``: level axis
    stored model top first (``pres_lev_pos`` still counts from the surface)."""
    L = _lib.lib()
    nlev, ngrid = u_2d.shape
    if type_grid not in (1, 2):
        raise ValueError('type_grid must be 1 (model levels) or 2 (pressure levels)')
    f3, f1, _, dt, layout, mem, ref = A.prepare_fields([u_2d, v_2d, aglh_2d], [u_s, v_s, aglh_s])
    if any(tuple(a.shape) != (nlev, ngrid) for a in f3) or any(a.shape[0] != ngrid for a in f1):
        raise ValueError('Input arrays must have the same shape.')
    u_, v_, h_ = f3
    us_, vs_, hs_ = f1
    start = None
    if type_grid == 2 and pres_lev_pos is not None:
        if A.is_cuda(ref):
            import torch
            start = torch.as_tensor(pres_lev_pos, device=ref.device).to(torch.int32).expand(ngrid).contiguous()
        else:
            start = np.ascontiguousarray(np.broadcast_to(np.asarray(pres_lev_pos), (ngrid,)), dtype=np.int32)
    want_all = (output != 1)
    srm = A.empty_like_host_or_device(ref, (ngrid,), 'float64')
    slm = A.empty_like_host_or_device(ref, (ngrid,), 'float64')
    rm = A.empty_like_host_or_device(ref, (ngrid, 2), 'float32') if want_all else None
    lm = A.empty_like_host_or_device(ref, (ngrid, 2), 'float32') if want_all else None
    m6 = A.empty_like_host_or_device(ref, (ngrid, 2), 'float32') if want_all else None
    A.order_on_stream(ref, stream, [u_, v_, h_, us_, vs_, hs_, start, srm, slm, rm, lm, m6])
    rc = L.xcape_cuda_srh_from_heights(A.ptr(u_), A.ptr(v_), A.ptr(h_), A.ptr(us_), A.ptr(vs_), A.ptr(hs_),
                                       C.c_int64(ngrid), nlev, dt, layout | (_lib.LEVELS_TOP_FIRST if top_first else 0), mem,
                                       C.c_double(float(depth)), A.ptr(start),
                                       A.ptr(srm), A.ptr(slm), A.ptr(rm), A.ptr(lm), A.ptr(m6),
                                       A.device_of(ref, device), A.stream_of(ref, stream))
    _lib.check(rc)
    if not want_all:
        return srm, slm
    tr = (lambda a: a.t()) if A.is_cuda(ref) else (lambda a: a.T)
    return srm, slm, tr(rm), tr(lm), tr(m6)
