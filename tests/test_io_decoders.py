"""xcape_b200.io — zarr-v2 and NetCDF-3 chunk decoders in front of the streamed executor.

The files are written here by small stand-alone writers (neither `zarr` nor `netCDF4` is installed); where scipy is
importable its NetCDF-3 reader is the referee for the writer, so that the reader under test is not checked against
itself only.  CPU tests cover the decoders and the step loaders (method='dummy'); the `gpu` tests stream an
archive through `stream_cape` / `stream_srh` and compare with direct calls.
"""
import bz2
import gzip
import json
import lzma
import os
import struct
import zlib

import numpy as np
import pytest

from xcape_b200 import io as xio


# ------------------------------------------------------------------------------------------------ writers
def write_zarr(path, a, chunks, compressor=None, sep='.', order='C', attrs=None):
    os.makedirs(path)
    meta = dict(zarr_format=2, shape=list(a.shape), chunks=list(chunks), dtype=a.dtype.str, compressor=compressor,
                fill_value=None, order=order, filters=None)
    if sep != '.':
        meta['dimension_separator'] = sep
    with open(os.path.join(path, '.zarray'), 'w') as f:
        json.dump(meta, f)
    if attrs:
        with open(os.path.join(path, '.zattrs'), 'w') as f:
            json.dump(attrs, f)
    enc = {None: lambda b: b, 'zlib': zlib.compress, 'gzip': gzip.compress, 'bz2': bz2.compress, 'lzma': lzma.compress}[
        compressor['id'] if compressor else None]
    grid = [range((n + c - 1) // c) for n, c in zip(a.shape, chunks)]
    for idx in np.ndindex(*[len(g) for g in grid]):
        block = np.zeros(chunks, dtype=a.dtype)      # edge chunks are stored at full chunk size
        sl = tuple(slice(i * c, min((i + 1) * c, n)) for i, c, n in zip(idx, chunks, a.shape))
        block[tuple(slice(0, s.stop - s.start) for s in sl)] = a[sl]
        name = sep.join(map(str, idx))
        fn = os.path.join(path, *name.split('/'))
        os.makedirs(os.path.dirname(fn), exist_ok=True)
        with open(fn, 'wb') as f:
            f.write(enc(block.tobytes(order=order)))


_NC = {'i2': (3, '>i2'), 'i4': (4, '>i4'), 'f4': (5, '>f4'), 'f8': (6, '>f8')}


def write_netcdf3(path, dims, variables, offset64=False):
    """dims: [(name, length or None for the record dimension)]; variables: [(name, dim names, numpy array, attrs)]."""
    def pad(b):
        return b + b'\x00' * (-len(b) % 4)

    def name(s):
        return struct.pack('>i', len(s)) + pad(s.encode())

    def attrs(d):
        if not d:
            return struct.pack('>ii', 0, 0)
        out = struct.pack('>ii', 12, len(d))
        for k, v in d.items():
            if isinstance(v, str):
                out += name(k) + struct.pack('>ii', 2, len(v)) + pad(v.encode())
            else:
                v = np.atleast_1d(np.asarray(v))
                t, dt = _NC[v.dtype.kind + str(v.dtype.itemsize)]
                out += name(k) + struct.pack('>ii', t, v.size) + pad(v.astype(dt).tobytes())
        return out

    dimid = {n: i for i, (n, _) in enumerate(dims)}
    rec = [n for n, ln in dims if ln is None]
    numrecs = max([a.shape[0] for _, dn, a, _ in variables if dn and dn[0] in rec] or [0])
    head = b'CDF' + (b'\x02' if offset64 else b'\x01') + struct.pack('>i', numrecs)
    head += struct.pack('>ii', 10, len(dims)) + b''.join(name(n) + struct.pack('>i', 0 if ln is None else ln) for n, ln in dims)
    head += attrs({'title': 'xcape_b200 test file'})
    recvars = [v for v in variables if v[1] and v[1][0] in rec]
    sizes = []
    for nm, dn, a, at in variables:
        t, dt = _NC[a.dtype.kind + str(a.dtype.itemsize)]
        per = int(np.prod(a.shape[1:] if (dn and dn[0] in rec) else a.shape, dtype=np.int64)) * np.dtype(dt).itemsize
        sizes.append(per if len(recvars) == 1 and dn and dn[0] in rec else (per + 3) // 4 * 4)
    osz = 8 if offset64 else 4
    var_head = sum(len(name(nm)) + 4 + 4 * len(dn) + len(attrs(at)) + 4 + 4 + osz for nm, dn, a, at in variables)
    begin = len(head) + 8 + var_head
    fixed = [i for i, v in enumerate(variables) if not (v[1] and v[1][0] in rec)]
    begins = {}
    for i in fixed:
        begins[i] = begin
        begin += sizes[i]
    rec_start = begin
    for i, v in enumerate(variables):
        if i not in fixed:
            begins[i] = begin
            begin += sizes[i]
    recsize = sum(sizes[i] for i in range(len(variables)) if i not in fixed)
    out = head + struct.pack('>ii', 11, len(variables))
    for i, (nm, dn, a, at) in enumerate(variables):
        t, dt = _NC[a.dtype.kind + str(a.dtype.itemsize)]
        out += name(nm) + struct.pack('>i', len(dn)) + b''.join(struct.pack('>i', dimid[d]) for d in dn) + attrs(at)
        out += struct.pack('>ii', t, sizes[i]) + struct.pack('>q' if offset64 else '>i', begins[i])
    assert len(out) == len(head) + 8 + var_head
    body = bytearray(rec_start - len(out) + recsize * numrecs)
    for i in fixed:
        nm, dn, a, at = variables[i]
        raw = a.astype(_NC[a.dtype.kind + str(a.dtype.itemsize)][1]).tobytes()
        body[begins[i] - len(out):begins[i] - len(out) + len(raw)] = raw
    for k in range(numrecs):
        for i, (nm, dn, a, at) in enumerate(variables):
            if i in fixed:
                continue
            raw = a[k].astype(_NC[a.dtype.kind + str(a.dtype.itemsize)][1]).tobytes()
            o = begins[i] - len(out) + k * recsize
            body[o:o + len(raw)] = raw
    with open(path, 'wb') as f:
        f.write(out + bytes(body))


# ------------------------------------------------------------------------------------------------ zarr
@pytest.mark.parametrize('compressor', [None, {'id': 'zlib', 'level': 1}, {'id': 'gzip', 'level': 1}, {'id': 'bz2', 'level': 1},
                                        {'id': 'lzma'}])
@pytest.mark.parametrize('sep,order', [('.', 'C'), ('/', 'C'), ('.', 'F')])
def test_zarr_v2_array_roundtrip(tmp_path, compressor, sep, order):
    rng = np.random.default_rng(0)
    a = rng.normal(size=(5, 7, 6, 11)).astype('<f4')
    write_zarr(str(tmp_path / 't'), a, (1, 3, 4, 5), compressor, sep, order, attrs={'units': 'K'})
    z = xio.ZarrArray(str(tmp_path / 't'))
    assert z.shape == a.shape and z.chunks == (1, 3, 4, 5) and z.dtype == a.dtype and z.attrs['units'] == 'K'
    assert np.array_equal(z[:], a) and np.array_equal(z[3], a[3]) and np.array_equal(z[-1, 2:6], a[-1, 2:6])
    assert np.array_equal(z[1:4, 1, 2:5, 3:10], a[1:4, 1, 2:5, 3:10])
    with pytest.raises(IndexError):
        z[5]


def test_zarr_unsupported_codec_and_missing_chunk(tmp_path):
    a = np.arange(24, dtype='<i2').reshape(2, 3, 4)
    write_zarr(str(tmp_path / 'a'), a, (1, 3, 4))
    os.remove(str(tmp_path / 'a' / '1.0.0'))          # an unwritten chunk reads as fill
    z = xio.ZarrArray(str(tmp_path / 'a'))
    assert np.array_equal(z[0], a[0]) and (z[1] == 0).all()
    meta = json.load(open(tmp_path / 'a' / '.zarray'))
    meta['compressor'] = {'id': 'blosc', 'cname': 'lz4'}
    json.dump(meta, open(tmp_path / 'a' / '.zarray', 'w'))
    with pytest.raises(NotImplementedError):
        xio.ZarrArray(str(tmp_path / 'a'))


# ------------------------------------------------------------------------------------------------ NetCDF-3
@pytest.mark.parametrize('offset64', [False, True])
@pytest.mark.parametrize('single_record_var', [False, True])
def test_netcdf3_reader(tmp_path, offset64, single_record_var):
    rng = np.random.default_rng(1)
    lev = np.array([1000., 850., 500.], dtype='f4')
    t = (rng.normal(size=(4, 3, 5, 6)) * 10 + 270).astype('f4')
    packed = rng.integers(-30000, 30000, size=(4, 5, 6)).astype('i2')
    packed[0, 0, 0] = -32767
    variables = [('level', ('level',), lev, {'units': 'millibars'}),
                 ('t', ('time', 'level', 'lat', 'lon'), t, {'units': 'K'})]
    if not single_record_var:
        variables.append(('sp', ('time', 'lat', 'lon'), packed,
                          {'scale_factor': np.float64(0.5), 'add_offset': np.float64(90000.0), '_FillValue': np.int16(-32767)}))
    fn = str(tmp_path / 'a.nc')
    write_netcdf3(fn, [('time', None), ('level', 3), ('lat', 5), ('lon', 6)], variables, offset64)
    try:                                               # referee for the writer: scipy's reader, where available
        from scipy.io import netcdf_file
        with netcdf_file(fn, 'r', mmap=False, maskandscale=False) as ref:
            assert np.array_equal(ref.variables['t'][:], t) and np.array_equal(ref.variables['level'][:], lev)
            if not single_record_var:
                assert np.array_equal(ref.variables['sp'][:], packed)
    except ImportError:
        pass
    f = xio.NetCDF3File(fn)
    assert f.dimensions == {'time': 4, 'level': 3, 'lat': 5, 'lon': 6} and f.attrs['title'] == 'xcape_b200 test file'
    v = f.variables['t']
    assert v.dims == ('time', 'level', 'lat', 'lon') and v.shape == t.shape and v.attrs['units'] == 'K'
    for k in range(4):
        assert np.array_equal(v[k], t[k]) and v[k].dtype == np.float32
    assert np.array_equal(f.variables['level'][:], lev)
    if not single_record_var:
        sp = f.variables['sp'][2]
        assert sp.dtype == np.float32 and np.array_equal(sp, packed[2].astype(np.float32) * np.float32(0.5) + np.float32(90000.0))
        assert np.isnan(f.variables['sp'][0][0, 0])
    with open(str(tmp_path / 'h5.nc'), 'wb') as g:
        g.write(b'\x89HDF\r\n\x1a\n' + b'\x00' * 64)
    with pytest.raises(ValueError):
        xio.NetCDF3File(str(tmp_path / 'h5.nc'))


# ------------------------------------------------------------------------------------------------ step loaders
def _archive(tmp_path, nt=3, nlev=37, ny=8, nx=16, fmt='zarr'):
    """A tiny ERA5-like pressure-level archive (K, Pa) from the synthetic soundings, as zarr or NetCDF-3."""
    from xcape_b200.synthetic import make_soundings
    steps = [make_soundings('C2', cols=(k * 5000, k * 5000 + ny * nx)) for k in range(nt)]
    lm = lambda k: np.stack([s[k].reshape(ny, nx, nlev).transpose(2, 0, 1) for s in steps])      # [time, level, lat, lon]
    sf = lambda k: np.stack([s[k].reshape(ny, nx) for s in steps])
    arch = dict(t=lm('t') + np.float32(273.15), td=lm('td') + np.float32(273.15), u=lm('u'), v=lm('v'),
                sp=sf('ps') * np.float32(100.0), t2m=sf('ts') + np.float32(273.15), d2m=sf('tds') + np.float32(273.15),
                u10=sf('us'), v10=sf('vs'))
    arch = {k: a.astype('<f4') for k, a in arch.items()}
    if fmt == 'zarr':
        for k, a in arch.items():
            write_zarr(str(tmp_path / k), a, (1, 8) + a.shape[2:] if a.ndim == 4 else (1,) + a.shape[1:], {'id': 'zlib', 'level': 1})
        opened = {k: xio.ZarrArray(str(tmp_path / k)) for k in arch}
    else:
        dims = [('time', None), ('level', nlev), ('latitude', ny), ('longitude', nx)]
        vs = [(k, ('time', 'level', 'latitude', 'longitude') if a.ndim == 4 else ('time', 'latitude', 'longitude'), a, {}) for k, a in arch.items()]
        write_netcdf3(str(tmp_path / 'era5.nc'), dims, vs, offset64=True)
        opened = xio.NetCDF3File(str(tmp_path / 'era5.nc')).variables
    return steps, opened, steps[0]['p']


@pytest.mark.parametrize('fmt', ['zarr', 'netcdf3'])
def test_step_loaders_convert_units_and_feed_the_streamed_executor(tmp_path, fmt):
    from xcape_b200 import stream
    steps, A, lev = _archive(tmp_path, fmt=fmt)
    loaders = xio.cape_steps(A['t'], A['td'], A['sp'], A['t2m'], A['d2m'], p_levels=lev)
    assert len(loaders) == 3
    p, t, td, ps, ts, tds = loaders[1]()
    s = steps[1]
    assert t.shape == (37, 8, 16) and t.dtype == np.float32 and t.flags['C_CONTIGUOUS']
    assert np.allclose(t, s['t'].reshape(8, 16, 37).transpose(2, 0, 1), atol=2e-5)       # K -> degC in float32
    assert np.allclose(ps, s['ps'].reshape(8, 16), rtol=1e-6) and np.array_equal(p, lev)
    out = list(stream.stream_cape(loaders, lev_axis=0, method='dummy', vertical_lev='pressure', source='most-unstable', pinned=False))
    assert len(out) == 3 and all(len(o) == 4 and o[0].shape == (8, 16) for o in out)
    sl = xio.srh_steps(A['t'], A['td'], A['u'], A['v'], A['sp'], A['t2m'], A['d2m'], A['u10'], A['v10'], p_levels=lev, times=[2, 0])
    assert len(sl) == 2 and len(sl[0]()) == 10 and np.array_equal(sl[0]()[3], steps[2]['u'].reshape(8, 16, 37).transpose(2, 0, 1))
    with pytest.raises(ValueError):
        xio.cape_steps(A['t'], A['td'], A['sp'], A['t2m'], A['d2m'])


@pytest.mark.gpu
@pytest.mark.parametrize('fmt', ['zarr', 'netcdf3'])
def test_streamed_archive_matches_direct_calls_on_the_gpu(tmp_path, fmt):
    from xcape_b200 import core, stream
    steps, A, lev = _archive(tmp_path, nt=4, fmt=fmt)
    kw = dict(source='most-unstable', pinc=500., vertical_lev='pressure')
    got = list(stream.stream_cape(xio.cape_steps(A['t'], A['td'], A['sp'], A['t2m'], A['d2m'], p_levels=lev), lev_axis=0,
                                  readers=2, **kw))
    for s, g in zip(steps, got):
        # the archive stores K / Pa in float32: converting back costs an ulp or two on the inputs, so compare with the
        # direct call on the SAME converted inputs rather than on the original degC / hPa fields
        t = (s['t'] + np.float32(273.15)) - np.float32(273.15)
        td = (s['td'] + np.float32(273.15)) - np.float32(273.15)
        ps = (s['ps'] * np.float32(100.0)) * np.float32(0.01)
        ts = (s['ts'] + np.float32(273.15)) - np.float32(273.15)
        tds = (s['tds'] + np.float32(273.15)) - np.float32(273.15)
        ref = core.calc_cape(lev, t, td, ps, ts, tds, **kw)
        for a, b in zip(g, ref):
            assert np.array_equal(np.asarray(a).ravel(), b)
    sg = list(stream.stream_srh(xio.srh_steps(A['t'], A['td'], A['u'], A['v'], A['sp'], A['t2m'], A['d2m'], A['u10'], A['v10'],
                                              p_levels=lev), lev_axis=0, depth=3000, vertical_lev='pressure'))
    assert len(sg) == 4 and all(np.isfinite(x[0]).all() for x in sg)
