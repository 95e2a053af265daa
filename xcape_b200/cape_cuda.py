"""CUDA counterpart of the reference's ``cape_fortran.py`` shim.

``cape(...)`` has the 13-argument signature and the 4-tuple return of
``xcape.cape_fortran.cape`` (cape_fortran.py:3-69) so that ``core._calc_cape_numpy`` can
dispatch ``method='cuda'`` beside ``'fortran'`` (core.py:313-324).  Where the reference calls
the f2py routines ``loopcape_ml`` / ``loopcape_pl1d`` this calls ``xcape_cuda_cape`` through
ctypes (include/xcape_b200.h).
"""
import ctypes as C

import numpy as np

from . import _array as A
from . import _lib


def _wider_than_fields(one_d, fields):
    """True if a 1-D argument is not float32 while every 3-D field is (the mixed-dtype case of ERA5 inputs)."""
    name = lambda a: str(a.dtype).replace('torch.', '')
    return all(name(f) == 'float32' for f in fields) and any(name(a) != 'float32' for a in one_d)


def cape(p_2d, t_2d, td_2d, p_s, t_s, td_s, flag_1d, pres_lev_pos, source, ml_depth, adiabat, pinc,
         type_grid, *, device=0, devices=None, stream=None, precision='faithful', return_status=False,
         return_counters=False, top_first=False):
    """
    Parameters follow ``cape_fortran.cape`` (cape_fortran.py:5-46): ``*_2d`` are
    ``(nlev, ngrid)`` (``p_2d`` is ``(nlev, 1)`` / ``(nlev,)`` when ``flag_1d == 1``), ``*_s`` are
    ``(ngrid,)``; ``source`` 1/2/3, ``adiabat`` 1..4, ``type_grid`` 1 (model) / 2 (pressure).

    ``pres_lev_pos``: 1-based first level with ``p <= ps`` per column (core.py:286-289), or
    ``None`` to have it computed on the device (pressure grids only).

    Extra keyword-only arguments: ``device`` (CUDA ordinal for host inputs), ``devices`` (list:
    shard host inputs in contiguous column blocks over several GPUs), ``stream`` (raw
    ``cudaStream_t`` for device inputs; default torch's current stream), ``precision``
    (``'faithful'``: bit-identical to the oracle's SPEC arithmetic; ``'fast'``: FP32-pipe moist
    iteration, tolerance-level parity, MU level still exact), ``return_status``
    (append the per-column status word), ``return_counters`` (append status and the number of
    moist-adiabat iterations each column ran), ``top_first`` (the level axis is stored model top
    first, as in ERA5 downloads; it is walked backwards on the device.  ``pres_lev_pos`` in and
    ``MUlev`` out keep counting from the surface).

    Returns ``CAPE, CIN, MUlev, zMUlev`` (float32, float32, int32, float32; shape ``(ngrid,)``)
    — numpy arrays for host inputs, torch CUDA tensors for CUDA-tensor inputs.
    """
    L = _lib.lib()
    if precision not in _lib.PRECISION:
        raise ValueError(f"precision must be one of {list(_lib.PRECISION)}")
    prec = _lib.PRECISION[precision]
    nlev, ngrid = t_2d.shape
    if type_grid == 1:
        p_is_1d = 0
    elif type_grid == 2:
        if flag_1d != 1:
            # the reference leaves CAPE undefined here (cape_fortran.py:54-67 -> NameError)
            raise ValueError('pressure-level grids need a 1-D pressure array (flag_1d == 1)')
        p_is_1d = 1
    else:
        raise ValueError('type_grid must be 1 (model levels) or 2 (pressure levels)')

    # The reference's routine is single precision and f2py casts each argument to float32 separately
    # (SURVEY §8b "Ownership"), so the working dtype follows the 3-D fields; 1-D arguments are cast to it.
    if p_is_1d:
        if pres_lev_pos is None and _wider_than_fields([p_2d, p_s], [t_2d, td_2d]):
            # core.py:286-289 evaluates ps - p in the arrays' own (wider) dtype: keep that for the start levels
            p_up = p_2d.reshape(-1)
            if top_first:                         # start levels count from the surface whatever the storage order
                p_up = p_up.flip(0) if A.is_torch(p_up) else p_up[::-1]
            pres_lev_pos = _device_pres_lev_pos(p_up, p_s, device=A.device_of(t_2d, devices[0] if devices else device),
                                                stream=stream)
        f3, f1, p, dt, layout, mem, ref = A.prepare_fields([t_2d, td_2d], [p_s, t_s, td_s], p=p_2d, dtype_from='fields')
        t_, td_ = f3
        if p.shape[0] != nlev:
            raise ValueError('p must have nlev entries')
    else:
        if tuple(p_2d.shape) != (nlev, ngrid):
            raise ValueError('p_2d must have the shape of t_2d on model levels')
        f3, f1, _, dt, layout, mem, ref = A.prepare_fields([p_2d, t_2d, td_2d], [p_s, t_s, td_s], dtype_from='fields')
        p, t_, td_ = f3
    ps_, ts_, tds_ = f1
    if tuple(td_.shape) != (nlev, ngrid) or any(a.shape[0] != ngrid for a in f1):
        raise ValueError('Input arrays must have the same shape.')

    start = None
    if p_is_1d and pres_lev_pos is not None:
        if A.is_cuda(ref):
            import torch
            start = torch.as_tensor(pres_lev_pos, device=ref.device).to(torch.int32).expand(ngrid).contiguous()
        else:
            start = np.ascontiguousarray(np.broadcast_to(np.asarray(pres_lev_pos), (ngrid,)), dtype=np.int32)

    cape_o = A.empty_like_host_or_device(ref, (ngrid,), 'float32')
    cin_o = A.empty_like_host_or_device(ref, (ngrid,), 'float32')
    mu_o = A.empty_like_host_or_device(ref, (ngrid,), 'int32')
    z_o = A.empty_like_host_or_device(ref, (ngrid,), 'float32')
    st_o = A.empty_like_host_or_device(ref, (ngrid,), 'int32') if (return_status or return_counters) else None
    it_o = A.empty_like_host_or_device(ref, (ngrid,), 'int32') if return_counters else None

    A.order_on_stream(ref, stream, [p, t_, td_, ps_, ts_, tds_, start, cape_o, cin_o, mu_o, z_o, st_o, it_o])

    def call(c0, c1, dev):
        n = c1 - c0
        if n <= 0:
            return

        es = 4 if dt == _lib.F32 else 8
        whole = (c0 == 0 and c1 == ngrid)
        keep = []                                  # dense per-shard copies of level-major fields stay alive for the call

        def off3(a):
            if layout == _lib.LEVEL_LAST:
                return A.ptr(a) + c0 * nlev * es
            if whole:
                return A.ptr(a)
            # a column block of a level-major host array is strided: hand the shard over as its own
            # dense [nlev, n] block (the C ABI takes the block width as the row pitch)
            keep.append(np.ascontiguousarray(a[:, c0:c1]))
            return A.ptr(keep[-1])

        def off1(a, es):
            return None if a is None else A.ptr(a) + c0 * es

        rc = L.xcape_cuda_cape(
            A.ptr(p) if p_is_1d else off3(p), off3(t_), off3(td_), off1(ps_, es), off1(ts_, es), off1(tds_, es),
            C.c_int64(n), nlev, p_is_1d, dt, layout | (_lib.LEVELS_TOP_FIRST if top_first else 0), mem, int(source), int(adiabat),
            C.c_float(float(ml_depth)), C.c_float(float(pinc)), off1(start, 4),
            off1(cape_o, 4), off1(cin_o, 4), off1(mu_o, 4), off1(z_o, 4), off1(st_o, 4), off1(it_o, 4),
            prec, dev, A.stream_of(ref, stream))
        _lib.check(rc)

    if mem == _lib.MEM_HOST and devices is not None and len(devices) > 1:
        # one C call: the library splits [0, ngrid) into contiguous 128-aligned blocks (the rule of
        # sharding.column_blocks) and runs each block's host path on its device from its own host thread
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        rc = L.xcape_cuda_cape_multi(
            A.ptr(p), A.ptr(t_), A.ptr(td_), A.ptr(ps_), A.ptr(ts_), A.ptr(tds_), C.c_int64(ngrid), nlev, p_is_1d, dt,
            layout | (_lib.LEVELS_TOP_FIRST if top_first else 0), int(source), int(adiabat),
            C.c_float(float(ml_depth)), C.c_float(float(pinc)), None if start is None else A.ptr(start),
            A.ptr(cape_o), A.ptr(cin_o), A.ptr(mu_o), A.ptr(z_o), None if st_o is None else A.ptr(st_o),
            None if it_o is None else A.ptr(it_o), prec, devs, len(devices))
        _lib.check(rc)
    else:
        dev = A.device_of(ref, devices[0] if devices else device)
        call(0, ngrid, dev)

    if return_counters:
        return cape_o, cin_o, mu_o, z_o, st_o, it_o
    if return_status:
        return cape_o, cin_o, mu_o, z_o, st_o
    return cape_o, cin_o, mu_o, z_o


def pres_lev_pos(p_1d, p_s, *, device=0, stream=None):
    """Device version of core.py:286-289: 1-based index of the first level with p <= ps
    (1 when every level lies below the surface), evaluated in the inputs' dtype."""
    L = _lib.lib()
    on_dev = [A.is_cuda(p_1d), A.is_cuda(p_s)]
    if any(on_dev) and not all(on_dev):
        raise ValueError('inputs must all be host arrays or all CUDA tensors')
    if not any(on_dev):
        p_1d, p_s = A.to_host_numpy(p_1d), A.to_host_numpy(p_s)
    dt = A.common_dtype([p_1d, p_s])
    p = A.dense_1d(A.cast(p_1d, dt))
    ps_ = A.dense_1d(A.cast(p_s, dt))
    n = ps_.shape[0]
    out = A.empty_like_host_or_device(ps_, (n,), 'int32')
    A.order_on_stream(ps_, stream, [p, ps_, out])
    rc = L.xcape_cuda_pres_lev_pos(A.ptr(p), A.ptr(ps_), C.c_int64(n), int(p.shape[0]),
                                   _lib.F32 if dt == 'float32' else _lib.F64,
                                   _lib.MEM_DEVICE if all(on_dev) else _lib.MEM_HOST, A.ptr(out),
                                   A.device_of(ps_, device), A.stream_of(ps_, stream))
    _lib.check(rc)
    return out


_device_pres_lev_pos = pres_lev_pos     # cape() has a parameter of the same name
