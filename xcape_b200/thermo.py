"""Thermodynamic helpers upstream of the column kernels.

``dewpoint_from_q`` turns the specific humidity that ERA5-style archives ship into the dew point
``calc_cape`` / ``calc_srh`` take — on the device, in one pass, through ``xcape_cuda_dewpoint_from_q``.
The reference has no such routine (doc/tutorial.rst:19-23 points at an external script), so this op is
"parity unpinned": it implements the inverse of the kernels' own Bolton saturation law
(CAPE_CODE_model_lev.f90:570-581) and is tested against that formula in float64.
"""
import ctypes as C

import numpy as np

from . import _array as A
from . import _lib


def dewpoint_from_q(p, q, *, lev_axis=-1, q_min=1e-10, device=0, stream=None):
    """Dew point (degC) from pressure ``p`` (hPa) and specific humidity ``q`` (kg/kg).

    ``q``: ``[..., nlev]`` (or ``[nlev, ...]`` with ``lev_axis=0``); ``p``: same shape, or 1-D
    ``[nlev]``.  ``q`` below ``q_min`` is raised to ``q_min`` first (``q_min=0`` keeps it: ``q <= 0``
    then gives NaN / -inf as the formula does).  Returns an array of ``q``'s shape, dtype and memory
    space (numpy in -> numpy out, CUDA tensor in -> CUDA tensor out); float32 stays float32, anything
    else is computed as float64.
    """
    L = _lib.lib()
    on_dev = [A.is_cuda(p), A.is_cuda(q)]
    if any(on_dev) and not all(on_dev):
        raise ValueError('inputs must all be host arrays or all CUDA tensors on one device')
    if not any(on_dev):
        p, q = A.to_host_numpy(p), A.to_host_numpy(q)
    if q.ndim < 1:
        raise ValueError('q needs a level axis')
    lev_axis = 0 if lev_axis == 0 else -1
    nlev = q.shape[lev_axis]
    p_is_1d = int(q.ndim > 1 and p.ndim == 1 and tuple(p.shape) == (nlev,))
    if not p_is_1d and tuple(p.shape) != tuple(q.shape):
        raise ValueError('p must have the shape of q or be 1-D [nlev]')
    dt = A.common_dtype([p, q])
    q_ = A.cast(q, dt)
    p_ = A.cast(p, dt)
    q_ = q_.contiguous() if A.is_torch(q_) else np.ascontiguousarray(q_)
    p_ = p_.contiguous() if A.is_torch(p_) else np.ascontiguousarray(p_)
    ncol = int(np.prod(q.shape)) // max(nlev, 1)
    layout = _lib.LEVEL_MAJOR if (lev_axis == 0 and q.ndim > 1) else _lib.LEVEL_LAST
    out = A.empty_like_host_or_device(q_, tuple(q.shape), dt)
    A.order_on_stream(q_, stream, [p_, q_, out])
    rc = L.xcape_cuda_dewpoint_from_q(A.ptr(p_), A.ptr(q_), C.c_int64(ncol), int(nlev), p_is_1d,
                                      _lib.F32 if dt == 'float32' else _lib.F64, layout,
                                      _lib.MEM_DEVICE if all(on_dev) else _lib.MEM_HOST, C.c_double(float(q_min)),
                                      A.ptr(out), A.device_of(q_, device), A.stream_of(q_, stream))
    _lib.check(rc)
    return out
