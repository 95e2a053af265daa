"""Array plumbing shared by the CUDA shims: numpy arrays (host memory) and torch tensors
(host, or CUDA device memory passed zero-copy as raw pointers).  torch is optional and is
only ever used for memory / streams — never for arithmetic on the product path."""
import os

import numpy as np

from . import _lib


POISON = -12345          # sentinel of XCAPE_B200_POISON_OUTPUTS (exact in float32 / float64 / int32)


def is_torch(a):
    return type(a).__module__.split('.')[0] == 'torch'


def is_cuda(a):
    return is_torch(a) and a.is_cuda


def to_host_numpy(a):
    """numpy view of a host array (torch CPU tensors are viewed, not copied)."""
    if is_torch(a):
        return a.detach().numpy()
    return np.asarray(a)


def common_dtype(arrs):
    """f2py casts each argument to the routine's dtype; we keep float32 when everything is
    float32 and otherwise promote everything to float64 (the kernels down-cast on load where
    the reference routine is single precision — same rounding as f2py's cast)."""
    names = {str(a.dtype).replace('torch.', '') for a in arrs}
    return 'float32' if names == {'float32'} else 'float64'


def cast(a, dtype):
    if is_torch(a):
        import torch
        return a.to(getattr(torch, dtype))
    return np.asarray(a, dtype=dtype)


def strides_elems(a):
    if is_torch(a):
        return tuple(a.stride())
    return tuple(s // a.itemsize for s in a.strides)


def layout_of_2d(a):
    """Layout of a (nlev, ncol) array, or None if neither dense form."""
    nlev, ncol = a.shape
    st = strides_elems(a)
    if (st == (ncol, 1)) or (nlev == 1 and st[1] == 1) or (ncol == 1 and st[0] == 1):
        return _lib.LEVEL_MAJOR
    if (st == (1, nlev)) or (ncol == 1 and st[0] == 1) or (nlev == 1 and st[1] == 1):
        return _lib.LEVEL_LAST
    return None


def force_layout(a, layout):
    """Dense copy of a (nlev, ncol) array in the requested layout (what f2py's
    Fortran-order coercion does for LEVEL_LAST)."""
    if is_torch(a):
        if layout == _lib.LEVEL_MAJOR:
            return a.contiguous()
        return a.t().contiguous().t()
    return np.ascontiguousarray(a) if layout == _lib.LEVEL_MAJOR else np.asfortranarray(a)


def dense_1d(a):
    if is_torch(a):
        return a.reshape(-1).contiguous()
    return np.ascontiguousarray(np.reshape(a, (-1,)))


def ptr(a):
    if a is None:
        return None
    if is_torch(a):
        return a.data_ptr()
    return a.ctypes.data


def empty_like_host_or_device(ref, shape, dtype):
    """Output buffer living where ``ref`` lives.  Uninitialised on purpose: the library writes every element
    of every output it is handed (gated, abandoned and work-list columns included), and zero-filling the four
    4 MB results of an ERA5 field cost 0.5 ms of a 14 ms call."""
    if is_cuda(ref):
        import torch
        return torch.empty(shape, dtype=getattr(torch, dtype), device=ref.device)
    if os.environ.get('XCAPE_B200_POISON_OUTPUTS'):       # tests: prove that nothing is left unwritten
        return np.full(shape, POISON, dtype=dtype)
    return np.empty(shape, dtype=dtype)


def stream_of(ref, stream=None):
    if stream is not None:
        return int(stream)
    if is_cuda(ref):
        import torch
        return int(torch.cuda.current_stream(ref.device).cuda_stream)
    return None


def order_on_stream(ref, stream, tensors):
    """An explicit ``stream=`` with CUDA tensors: everything the shim prepared (casts, relayouts, start levels, the
    freshly allocated outputs) was produced on torch's CURRENT stream, while the kernels are enqueued on the
    caller's stream.  Make the caller's stream wait for the preparation, and tell torch's caching allocator that
    those tensors are in use on it — otherwise a temporary could be read before it is written, or be recycled
    while the kernel is still running."""
    if stream is None or not is_cuda(ref):
        return
    import torch
    cur = torch.cuda.current_stream(ref.device)
    if int(stream) == int(cur.cuda_stream):
        return
    ext = torch.cuda.ExternalStream(int(stream), device=ref.device)
    ext.wait_stream(cur)
    for t in tensors:
        if t is not None and is_cuda(t):
            t.record_stream(ext)


def device_of(ref, device):
    if is_cuda(ref):
        return ref.device.index if ref.device.index is not None else 0
    return int(device)


def prepare_fields(fields3d, fields1d, p=None, dtype_from='all'):
    """Bring a group of (nlev, ncol) fields + (ncol,) surface fields (+ optional 1-D p) to one
    dtype, one dense layout and one memory space.  Returns (f3, f1, p, dtype_code, layout,
    mem, ref) — copies are only made where f2py would have made them too.

    ``dtype_from='fields'``: the common dtype is taken from the 3-D fields alone and the per-column /
    per-level 1-D arrays are cast to it (``ncol`` resp. ``nlev`` elements) — what f2py does for a
    single-precision routine, which casts every argument separately.  A float64 ``ps`` or an integer
    / float64 pressure axis next to float32 fields (the usual ERA5 case: the ``level`` coordinate is
    int or float64) then no longer drags every 3-D field through a float64 host copy."""
    every = list(fields3d) + list(fields1d) + ([p] if p is not None else [])
    on_dev = [is_cuda(a) for a in every]
    if any(on_dev) and not all(on_dev):
        raise ValueError('inputs must all be host arrays or all CUDA tensors on one device')
    if not any(on_dev):
        every = [to_host_numpy(a) for a in every]
        n3, n1 = len(fields3d), len(fields1d)
        fields3d, fields1d = every[:n3], every[n3:n3 + n1]
        p = every[-1] if p is not None else None
    dt = common_dtype(every[:len(fields3d)] if dtype_from == 'fields' else every)
    fields3d = [cast(a, dt) for a in fields3d]
    fields1d = [dense_1d(cast(a, dt)) for a in fields1d]
    if p is not None:
        p = dense_1d(cast(p, dt))
    lays = {layout_of_2d(a) for a in fields3d}
    if len(lays) == 1 and None not in lays:
        layout = lays.pop()
    else:
        layout = _lib.LEVEL_LAST
        fields3d = [force_layout(a, layout) for a in fields3d]
    mem = _lib.MEM_DEVICE if all(on_dev) and on_dev else _lib.MEM_HOST
    return fields3d, fields1d, p, (_lib.F32 if dt == 'float32' else _lib.F64), layout, mem, fields3d[0]
