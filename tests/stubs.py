"""Minimal stand-ins for ``dask.array`` and ``xarray`` — TEST INFRASTRUCTURE.

Neither package is installed in the build image or on the GPU box (probed both), so the dask
(`core._calc_*_gufunc`) and xarray (`xcape_b200/xarray.py`) code paths would otherwise never execute.  These
stubs implement exactly the slice of the two APIs those paths use, with the semantics of the real functions:

* ``apply_gufunc(func, signature, *args, output_dtypes=, axis=-1, vectorize=False, **kwargs)``: parse the gufunc
  signature, broadcast the loop dimensions, call ``func`` once per BLOCK of the loop dimensions (core
  dimensions are never split) and stitch the per-block outputs together — dask's ``blockwise`` contract;
* ``xarray.apply_ufunc(func, *args, kwargs=, input_core_dims=, output_core_dims=, dask='parallelized', ...)``:
  move each argument's core dims last, align the remaining dims by NAME, call ``func`` on the bare arrays (through
  ``apply_gufunc`` when an argument is chunked), wrap the outputs as DataArrays on the broadcast dims.

``tests/test_dask_xarray_paths.py`` also runs against the real packages wherever they can be imported.
"""
import itertools
import re
import types

import numpy as np


# --------------------------------------------------------------------------------------------- dask.array
class Array:
    """A chunked array: numpy data + chunk sizes per axis.  ``compute()`` returns the numpy array."""

    def __init__(self, data, chunks):
        self._data = np.asarray(data)
        self.chunks = tuple(tuple(c) for c in chunks)
        assert tuple(sum(c) for c in self.chunks) == self._data.shape

    shape = property(lambda self: self._data.shape)
    ndim = property(lambda self: self._data.ndim)
    dtype = property(lambda self: self._data.dtype)

    def compute(self):
        return self._data

    def __array__(self, dtype=None, copy=None):
        raise TypeError('implicit conversion of a chunked array: the code under test must go through apply_gufunc')


def from_array(x, chunks):
    x = np.asarray(x)
    if isinstance(chunks, int):
        chunks = (chunks,) * x.ndim
    out = []
    for n, c in zip(x.shape, chunks):
        c = n if c in (-1, None) else c
        out.append(tuple([c] * (n // c) + ([n % c] if n % c else [])) if n else (0,))
    return Array(x, out)


def _parse(signature):
    ins, outs = signature.split('->')
    dims = lambda s: [tuple(d for d in g.split(',') if d) for g in re.findall(r'\(([^)]*)\)', s)]
    return dims(ins), dims(outs)


def apply_gufunc(func, signature, *args, output_dtypes=None, axis=-1, vectorize=False, **kwargs):
    assert axis == -1 and not vectorize
    in_core, out_core = _parse(signature)
    if len(in_core) != len(args):
        raise ValueError(f'signature {signature!r} has {len(in_core)} inputs, {len(args)} arguments given')
    if any(out_core):
        raise NotImplementedError('stub: outputs with core dimensions')
    arrs = [a if isinstance(a, Array) else from_array(a, -1) for a in args]
    loop_nd = max(a.ndim - len(c) for a, c in zip(arrs, in_core))
    loop_chunks = None
    for a, c in zip(arrs, in_core):
        nd = a.ndim - len(c)
        if len(c) and any(len(ch) != 1 for ch in a.chunks[nd:]):
            raise ValueError('core dimension is split over several chunks (dask would ask for a rechunk)')
        if nd == 0:
            continue                              # no loop dimensions: broadcast to every block
        if nd != loop_nd:
            raise NotImplementedError('stub: arguments with different numbers of loop dimensions')
        if loop_chunks is None:
            loop_chunks = a.chunks[:nd]
        elif a.chunks[:nd] != loop_chunks:
            raise ValueError('loop-dimension chunks of the arguments differ')
    loop_chunks = loop_chunks or ()
    loop_shape = tuple(sum(c) for c in loop_chunks)
    n_out = len(out_core)
    dts = list(output_dtypes) if isinstance(output_dtypes, (list, tuple)) else [output_dtypes] * n_out
    outs = [np.empty(loop_shape, dtype=np.dtype(d)) for d in dts]
    edges = [np.concatenate([[0], np.cumsum(c)]) for c in loop_chunks]
    calls = 0
    for idx in itertools.product(*[range(len(c)) for c in loop_chunks]):
        sl = tuple(slice(int(e[i]), int(e[i + 1])) for e, i in zip(edges, idx))
        blocks = []
        for a, c in zip(arrs, in_core):
            nd = a.ndim - len(c)
            blocks.append(a._data[(sl if nd else ()) + (slice(None),) * len(c)])
        res = func(*blocks, **kwargs)
        res = res if isinstance(res, tuple) else (res,)
        assert len(res) == n_out, f'function returned {len(res)} outputs, signature promises {n_out}'
        for o, r in zip(outs, res):
            o[sl] = np.asarray(r)            # dask casts to the declared output dtype
        calls += 1
    apply_gufunc.calls = calls
    wrapped = tuple(Array(o, loop_chunks) for o in outs)
    return wrapped if n_out > 1 else wrapped[0]


dask_array = types.ModuleType('dask.array')
dask_array.Array, dask_array.from_array, dask_array.apply_gufunc = Array, from_array, apply_gufunc


# ------------------------------------------------------------------------------------------------ xarray
class DataArray:
    def __init__(self, data, dims, attrs=None):
        self.data = data
        self.dims = tuple(dims)
        self.attrs = dict(attrs or {})
        assert len(self.dims) == data.ndim

    @property
    def values(self):
        return self.data.compute() if isinstance(self.data, Array) else np.asarray(self.data)

    @property
    def shape(self):
        return self.data.shape

    def chunk(self, sizes):
        return DataArray(from_array(self.values, tuple(sizes.get(d, -1) for d in self.dims)), self.dims, self.attrs)


class Dataset:
    def __init__(self, data_vars):
        self.data_vars = dict(data_vars)
        self.attrs = {}

    def __getitem__(self, k):
        return self.data_vars[k]

    def __contains__(self, k):
        return k in self.data_vars

    def __iter__(self):
        return iter(self.data_vars)


def apply_ufunc(func, *args, kwargs=None, input_core_dims=None, output_core_dims=((),), dask='forbidden',
                output_dtypes=None, dask_gufunc_kwargs=None):
    kwargs = kwargs or {}
    assert len(input_core_dims) == len(args)
    if any(len(c) for c in output_core_dims):
        raise NotImplementedError('stub: outputs with core dimensions')
    loop_dims, sizes = [], {}
    for a, core in zip(args, input_core_dims):
        for d, n in zip(a.dims, a.shape):
            if d not in core and d not in loop_dims:
                loop_dims.append(d)
            if sizes.setdefault(d, n) != n:
                raise ValueError(f'dimension {d!r} has conflicting sizes')
    bare = []
    for a, core in zip(args, input_core_dims):
        missing = [d for d in core if d not in a.dims]
        if missing:
            raise ValueError(f'core dimension {missing} not on an argument with dims {a.dims}')
        order = [d for d in loop_dims if d in a.dims] + list(core)
        perm = [a.dims.index(d) for d in order]
        x = a.data
        if isinstance(x, Array):
            x = Array(np.transpose(x._data, perm), [x.chunks[i] for i in perm])
        else:
            x = np.transpose(np.asarray(x), perm)
        have = [d for d in loop_dims if d in a.dims]
        if have != loop_dims and have:          # insert broadcast axes for loop dims this argument lacks
            shape = [sizes[d] if d in have else 1 for d in loop_dims] + [sizes[d] for d in core]
            x = x.reshape(shape) if not isinstance(x, Array) else x
        bare.append(x)
    chunked = any(isinstance(x, Array) for x in bare)
    if chunked:
        if dask != 'parallelized':
            raise ValueError("chunked input needs dask='parallelized'")
        sig = ','.join('(' + ','.join(c) + ')' for c in input_core_dims) + '->' + ','.join('()' for _ in output_core_dims)
        res = apply_gufunc(func, sig, *bare, output_dtypes=output_dtypes, **kwargs)
    else:
        res = func(*bare, **kwargs)
    res = res if isinstance(res, tuple) else (res,)
    assert len(res) == len(output_core_dims)
    out = tuple(DataArray(r, loop_dims) for r in res)
    return out if len(out) > 1 else out[0]


xarray = types.ModuleType('xarray')
xarray.DataArray, xarray.Dataset, xarray.apply_ufunc = DataArray, Dataset, apply_ufunc
