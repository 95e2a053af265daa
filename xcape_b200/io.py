"""Chunk decoders in front of :mod:`xcape_b200.stream`.

Reanalysis archives hold the 3-D fields as ``(time, level, lat, lon)`` arrays in zarr stores or NetCDF files; the
reference's tutorial (doc/tutorial.rst:19-36) leaves getting them into memory to xarray + dask.  Neither those
nor ``zarr`` / ``netCDF4`` / ``h5py`` exist in this image, and they are not needed for the two formats that are
plain files with a small header:

* **zarr v2** directory stores: ``.zarray`` is JSON (shape, chunks, dtype, compressor, order, fill_value,
  dimension_separator), every chunk is one file, compressed with nothing / zlib / gzip / bz2 / lzma — all in the
  Python standard library, all releasing the GIL while they decompress, so ``stream_cape``'s reader threads
  decode chunks in parallel.  (Blosc / zstd chunks need their codecs: a clear error says so.)
* **NetCDF-3** classic and 64-bit-offset files (what the Copernicus CDS served for ERA5 for years: ``short``
  variables with ``scale_factor`` / ``add_offset``): the header is parsed here, variables come back as
  ``numpy.memmap`` views straight onto the file, unpacked to float32 one time step at a time.

``cape_steps`` / ``srh_steps`` turn a set of such variables into the per-time-step loaders ``stream_cape`` /
``stream_srh`` take: each loader returns level-major ``[level, lat, lon]`` float32 fields (``lev_axis=0``: zero
relayout on the device), unit conversion (K -> degC, Pa -> hPa) and the optional q -> Td conversion included.
Nothing here touches the numerics of the column kernels.
"""
import bz2
import gzip
import json
import lzma
import os
import struct
import zlib

import numpy as np

__all__ = ['ZarrArray', 'NetCDF3File', 'cape_steps', 'srh_steps']


# ------------------------------------------------------------------------------------------------ zarr v2
def _decompressor(spec):
    if spec is None:
        return lambda b: b
    cid = spec.get('id')
    if cid == 'zlib':
        return zlib.decompress
    if cid == 'gzip':
        return gzip.decompress
    if cid == 'bz2':
        return bz2.decompress
    if cid == 'lzma':
        return lzma.decompress
    raise NotImplementedError(f'zarr compressor {cid!r} needs a codec that is not in the standard library '
                              '(supported: none, zlib, gzip, bz2, lzma)')


class ZarrArray:
    """One array of a zarr v2 directory store (read-only).  ``a[k]`` / ``a[k0:k1]`` / ``a[k, l0:l1]`` ... read the
    chunks that overlap the selection — basic indexing with integers and unit-step slices."""

    def __init__(self, path):
        self.path = path
        with open(os.path.join(path, '.zarray')) as f:
            meta = json.load(f)
        if meta.get('zarr_format') != 2:
            raise NotImplementedError('only zarr format 2 stores are supported')
        if meta.get('filters'):
            raise NotImplementedError('zarr filters are not supported')
        self.shape = tuple(meta['shape'])
        self.chunks = tuple(meta['chunks'])
        self.dtype = np.dtype(meta['dtype'])
        self.order = meta.get('order', 'C')
        self.fill_value = meta.get('fill_value')
        self.sep = meta.get('dimension_separator', '.')
        self._decode = _decompressor(meta.get('compressor'))
        self._raw = meta.get('compressor') is None
        self.attrs = {}
        zattrs = os.path.join(path, '.zattrs')
        if os.path.exists(zattrs):
            with open(zattrs) as f:
                self.attrs = json.load(f)
        self.ndim = len(self.shape)

    def _chunk(self, idx):
        """The decoded chunk with grid index ``idx`` (full chunk shape; edge chunks are stored padded)."""
        name = self.sep.join(str(i) for i in idx) if idx else '0'
        fn = os.path.join(self.path, *name.split('/')) if self.sep == '/' else os.path.join(self.path, name)
        if not os.path.exists(fn):                   # an unwritten chunk is all fill_value
            fv = 0 if self.fill_value in (None, 'NaN') and self.dtype.kind != 'f' else self.fill_value
            fv = np.nan if fv in (None, 'NaN') else fv
            return np.full(self.chunks, fv, dtype=self.dtype)
        with open(fn, 'rb') as f:
            raw = self._decode(f.read())
        return np.frombuffer(raw, dtype=self.dtype).reshape(self.chunks, order=self.order)

    def _read_into(self, cid, lo, hi, out):
        """Fast path: an uncompressed C-order chunk that lies wholly inside the selection and lands on a contiguous
        run of ``out`` is read from its file straight into place (no intermediate buffer, no second copy)."""
        if self._raw is False or self.order != 'C':
            return False
        dst = []
        for k, (c, l, h, n) in enumerate(zip(self.chunks, lo, hi, self.shape)):
            c0 = cid[k] * c
            if c0 < l or c0 + c > h:                 # chunk sticks out of the selection (or is a padded edge chunk)
                return False
            dst.append(slice(c0 - l, c0 - l + c))
        view = out[tuple(dst)]
        if not view.flags['C_CONTIGUOUS']:
            return False
        name = self.sep.join(str(i) for i in cid)
        fn = os.path.join(self.path, *name.split('/')) if self.sep == '/' else os.path.join(self.path, name)
        try:
            with open(fn, 'rb', buffering=0) as f:
                got = f.readinto(memoryview(view).cast('B'))
        except FileNotFoundError:
            return False
        if got != view.nbytes:
            raise IOError(f'short chunk file {fn}')
        return True

    def __getitem__(self, sel):
        sel = sel if isinstance(sel, tuple) else (sel,)
        if len(sel) > self.ndim:
            raise IndexError('too many indices')
        sel = sel + (slice(None),) * (self.ndim - len(sel))
        lo, hi, squeeze = [], [], []
        for s, n in zip(sel, self.shape):
            if isinstance(s, (int, np.integer)):
                s = int(s) + (n if s < 0 else 0)
                if not 0 <= s < n:
                    raise IndexError('index out of range')
                lo.append(s); hi.append(s + 1); squeeze.append(True)
            elif isinstance(s, slice):
                a, b, st = s.indices(n)
                if st != 1:
                    raise NotImplementedError('only unit-step slices')
                lo.append(a); hi.append(max(a, b)); squeeze.append(False)
            else:
                raise NotImplementedError('only integers and slices')
        out = np.empty([h - l for l, h in zip(lo, hi)], dtype=self.dtype)
        ranges = [range(l // c, (h - 1) // c + 1) if h > l else range(0) for l, h, c in zip(lo, hi, self.chunks)]
        for idx in np.ndindex(*[len(r) for r in ranges]):
            cid = tuple(r[i] for r, i in zip(ranges, idx))
            if self._read_into(cid, lo, hi, out):
                continue
            ch = self._chunk(cid)
            src, dst = [], []
            for k, (c, l, h) in enumerate(zip(self.chunks, lo, hi)):
                c0 = cid[k] * c
                a, b = max(l, c0), min(h, c0 + c)
                src.append(slice(a - c0, b - c0)); dst.append(slice(a - l, b - l))
            out[tuple(dst)] = ch[tuple(src)]
        return out.reshape([n for n, sq in zip(out.shape, squeeze) if not sq])


# ----------------------------------------------------------------------------------------------- NetCDF-3
_NC_TYPES = {1: ('i1', 1), 2: ('S1', 1), 3: ('>i2', 2), 4: ('>i4', 4), 5: ('>f4', 4), 6: ('>f8', 8)}


class _Reader:
    def __init__(self, buf, offset64):
        self.b, self.o, self.offset64 = buf, 0, offset64

    def i32(self):
        v = struct.unpack_from('>i', self.b, self.o)[0]
        self.o += 4
        return v

    def off(self):
        if self.offset64:
            v = struct.unpack_from('>q', self.b, self.o)[0]
            self.o += 8
            return v
        return self.i32()

    def name(self):
        n = self.i32()
        s = bytes(self.b[self.o:self.o + n]).decode()
        self.o += (n + 3) // 4 * 4
        return s

    def values(self, nc_type, n):
        dt, size = _NC_TYPES[nc_type]
        raw = bytes(self.b[self.o:self.o + n * size])
        self.o += (n * size + 3) // 4 * 4
        if nc_type == 2:
            return raw.decode(errors='replace').rstrip('\x00')
        v = np.frombuffer(raw, dtype=dt)
        return v[0].item() if n == 1 else v.astype(v.dtype.newbyteorder('='))

    def attrs(self):
        tag, n = self.i32(), self.i32()
        if tag == 0 and n == 0:
            return {}
        if tag != 12:
            raise ValueError('corrupt NetCDF header (attribute list)')
        out = {}
        for _ in range(n):
            k = self.name()
            t, m = self.i32(), self.i32()
            out[k] = self.values(t, m)
        return out


class NetCDF3Variable:
    """A variable of a NetCDF-3 file: ``.data`` is a (big-endian) ``numpy.memmap`` view onto the file, ``v[k]`` the
    k-th slab along the first axis unpacked to float32 (``scale_factor`` / ``add_offset``, ``_FillValue`` /
    ``missing_value`` -> NaN) — one time step per call, so an archive larger than memory streams through."""

    def __init__(self, name, dims, shape, attrs, data):
        self.name, self.dims, self.shape, self.attrs, self.data = name, dims, shape, attrs, data
        self.ndim = len(shape)

    def __getitem__(self, sel):
        raw = self.data[sel]
        a = np.asarray(raw)
        sf, ao = self.attrs.get('scale_factor'), self.attrs.get('add_offset')
        fill = self.attrs.get('_FillValue', self.attrs.get('missing_value'))
        if a.dtype.kind in 'iu' and (sf is not None or ao is not None):
            out = a.astype(np.float32) * np.float32(1.0 if sf is None else sf) + np.float32(0.0 if ao is None else ao)
            if fill is not None:
                out[a == fill] = np.nan
            return out
        out = a.astype(np.float32) if a.dtype.kind == 'f' else a.astype(a.dtype.newbyteorder('='))
        if fill is not None and out.dtype.kind == 'f':
            out[a == fill] = np.nan
        return out


class NetCDF3File:
    """NetCDF-3 classic (``CDF\\x01``) / 64-bit-offset (``CDF\\x02``) reader.  ``f.variables[name]`` ->
    :class:`NetCDF3Variable`; ``f.dimensions`` -> name -> length (the record dimension: number of records)."""

    def __init__(self, path):
        self.path = path
        self._mm = np.memmap(path, dtype=np.uint8, mode='r')
        b = self._mm
        if bytes(b[:3]) != b'CDF' or b[3] not in (1, 2):
            raise ValueError('not a NetCDF-3 classic / 64-bit-offset file (NetCDF-4 is HDF5 and needs h5py)')
        r = _Reader(b, offset64=(b[3] == 2))
        r.o = 4
        numrecs = r.i32()
        tag, n = r.i32(), r.i32()
        dims = []
        if tag == 10:
            for _ in range(n):
                nm = r.name()
                dims.append((nm, r.i32()))
        elif not (tag == 0 and n == 0):
            raise ValueError('corrupt NetCDF header (dimension list)')
        self.attrs = r.attrs()
        tag, n = r.i32(), r.i32()
        raw_vars = []
        if tag == 11:
            for _ in range(n):
                nm = r.name()
                nd = r.i32()
                dimids = [r.i32() for _ in range(nd)]
                at = r.attrs()
                t = r.i32()
                vsize = r.i32()
                begin = r.off()
                raw_vars.append((nm, dimids, at, t, vsize, begin))
        elif not (tag == 0 and n == 0):
            raise ValueError('corrupt NetCDF header (variable list)')
        rec_vars = [v for v in raw_vars if v[1] and dims[v[1][0]][1] == 0]
        recsize = sum(v[4] for v in rec_vars)
        if len(rec_vars) == 1:                       # a single record variable is stored unpadded
            nm, dimids, at, t, vsize, begin = rec_vars[0]
            recsize = int(np.prod([dims[i][1] for i in dimids[1:]], dtype=np.int64)) * _NC_TYPES[t][1]
        if numrecs < 0:                              # "streaming" record count: derive it from the file size
            first = min(v[5] for v in rec_vars) if rec_vars else 0
            numrecs = (b.size - first) // recsize if recsize else 0
        self.dimensions = {nm: (numrecs if ln == 0 else ln) for nm, ln in dims}
        self.variables = {}
        for nm, dimids, at, t, vsize, begin in raw_vars:
            dt, size = _NC_TYPES[t]
            dnames = tuple(dims[i][0] for i in dimids)
            if dimids and dims[dimids[0]][1] == 0:   # record variable: one slab per record, records interleaved
                inner = tuple(dims[i][1] for i in dimids[1:])
                cnt = int(np.prod(inner, dtype=np.int64))
                data = np.ndarray((numrecs,) + inner, dtype=dt, buffer=self._mm, offset=begin,
                                  strides=(recsize,) + tuple(int(np.prod(inner[k + 1:], dtype=np.int64)) * size for k in range(len(inner))))
                del cnt
                shape = (numrecs,) + inner
            else:
                shape = tuple(dims[i][1] for i in dimids)
                data = np.ndarray(shape, dtype=dt, buffer=self._mm, offset=begin)
            self.variables[nm] = NetCDF3Variable(nm, dnames, shape, at, data)

    def close(self):
        self.variables = {}
        del self._mm


# ---------------------------------------------------------------------------- steps for stream_cape / stream_srh
def _convert(a, kind):
    """K -> degC ('temperature'), Pa -> hPa ('pressure'), nothing (None); float32, dense."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    if kind == 'temperature':
        a -= np.float32(273.15)
    elif kind == 'pressure':
        a *= np.float32(0.01)
    return a


def _step_loader(k, fields3, fields2, p_axis, conv3, conv2, q_to_td, device):
    def load():
        f3 = [_convert(v[k], c) for v, c in zip(fields3, conv3)]                # [level, lat, lon] each
        f2 = [_convert(v[k], c) for v, c in zip(fields2, conv2)]                # [lat, lon] each
        if q_to_td is not None:                     # slot of the dew point holds specific humidity: convert on the GPU
            from .thermo import dewpoint_from_q
            i3, i2 = q_to_td
            p3 = p_axis if p_axis is not None else f3[0]
            f3[i3] = dewpoint_from_q(p3, f3[i3], lev_axis=0, device=device)
            f2[i2] = dewpoint_from_q(f2[0][None], f2[i2][None], lev_axis=0, device=device)[0]
        return ((p_axis,) if p_axis is not None else ()) + tuple(f3) + tuple(f2)
    return load


def cape_steps(t, td, ps, ts, tds, *, p=None, p_levels=None, kelvin=True, pascal=True, td_is_q=False, times=None, device=0):
    """Per-time-step loaders for :func:`xcape_b200.stream.stream_cape` (call it with ``lev_axis=0``).

    ``t, td`` (and ``p`` on model levels): ``[time, level, lat, lon]`` arrays with ``a[k]`` slab access
    (:class:`ZarrArray`, :class:`NetCDF3Variable`, ``numpy.memmap`` ...); ``ps, ts, tds``: ``[time, lat, lon]``.
    ``p_levels``: the 1-D pressure axis (hPa) of a pressure-level archive.  ``kelvin`` / ``pascal``: the archive
    stores K and Pa (ERA5 does) — converted to the reference's degC / hPa.  ``td_is_q``: the ``td`` / ``tds`` slots
    hold specific humidity (ERA5 ships q, not Td): converted with ``thermo.dewpoint_from_q`` on the GPU.
    """
    if (p is None) == (p_levels is None):
        raise ValueError('give either p (3-D, model levels) or p_levels (1-D pressure axis)')
    nt = t.shape[0]
    tk = 'temperature' if kelvin else None
    pk = 'pressure' if pascal else None
    f3 = ([p] if p is not None else []) + [t, td]
    c3 = ([pk] if p is not None else []) + [tk, None if td_is_q else tk]
    pax = None if p_levels is None else np.ascontiguousarray(p_levels, dtype=np.float32)
    q = (len(f3) - 1, 2) if td_is_q else None
    return [_step_loader(k, f3, [ps, ts, tds], pax, c3, [pk, tk, None if td_is_q else tk], q, device)
            for k in (range(nt) if times is None else times)]


def srh_steps(t, td, u, v, ps, ts, tds, us, vs, *, p=None, p_levels=None, kelvin=True, pascal=True, times=None, device=0):
    """As :func:`cape_steps` for :func:`xcape_b200.stream.stream_srh`: steps are ``(p, t, td, u, v, ps, ts, tds, us, vs)``."""
    if (p is None) == (p_levels is None):
        raise ValueError('give either p (3-D, model levels) or p_levels (1-D pressure axis)')
    nt = t.shape[0]
    tk = 'temperature' if kelvin else None
    pk = 'pressure' if pascal else None
    f3 = ([p] if p is not None else []) + [t, td, u, v]
    c3 = ([pk] if p is not None else []) + [tk, tk, None, None]
    pax = None if p_levels is None else np.ascontiguousarray(p_levels, dtype=np.float32)
    return [_step_loader(k, f3, [ps, ts, tds, us, vs], pax, c3, [pk, tk, tk, None, None], None, device)
            for k in (range(nt) if times is None else times)]
