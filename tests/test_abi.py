"""The C-ABI shared library loads on a CPU-only box and exports every symbol that
include/xcape_b200.h declares; argument validation works without touching a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from xcape_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, 'include', 'xcape_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(xcape_cuda_\w+)\s*\(', hdr))
    assert declared == set(lib.SYMBOLS)
    L = C.CDLL(lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name


def test_no_cpu_fallback_symbols(lib):
    """The product library must not link or embed the oracle."""
    import subprocess
    out = subprocess.run(['nm', '-D', '--defined-only', lib.LIB_PATH], capture_output=True, text=True).stdout
    assert 'xcape_ref_' not in out
    for mod in ('cape_cuda', 'srh_cuda', 'stdheight_cuda', 'core', '_lib', '_array', 'sharding'):
        src = open(os.path.join(ROOT, 'xcape_b200', mod + '.py')).read()
        assert 'import oracle' not in src and 'from oracle' not in src, mod


def test_version_and_error_codes(lib):
    L = lib.lib()
    assert b'sm_100a' in L.xcape_cuda_version()
    z = np.zeros(8, np.float32)
    zi = np.zeros(8, np.int32)
    p = lambda a: a.ctypes.data  # noqa: E731

    def cape(**over):
        kw = dict(ncol=8, nlev=1, p1d=0, dtype=0, layout=0, mem=0, source=1, adiabat=1, ml=500., pinc=500.)
        kw.update(over)
        return L.xcape_cuda_cape(p(z), p(z), p(z), p(z), p(z), p(z), C.c_int64(kw['ncol']), kw['nlev'], kw['p1d'],
                                 kw['dtype'], kw['layout'], kw['mem'], kw['source'], kw['adiabat'],
                                 C.c_float(kw['ml']), C.c_float(kw['pinc']), None, p(z), p(z), p(zi), p(z), None, None,
                                 0, 0, None)
    assert cape(source=4) == lib.ERR_ARG and b'source' in L.xcape_cuda_last_error()
    assert cape(adiabat=0) == lib.ERR_ARG
    assert cape(pinc=0.0) == lib.ERR_ARG
    assert cape(nlev=0) == lib.ERR_ARG
    assert cape(dtype=7) == lib.ERR_ARG
    assert cape(layout=2) == lib.ERR_ARG
    assert cape(ncol=-1) == lib.ERR_ARG
    assert cape(ncol=0) == lib.OK                      # empty grid: nothing to do, no device needed
    with pytest.raises(ValueError):
        lib.check(cape(source=0))


def test_multi_device_entries_validate_without_a_gpu(lib):
    """xcape_cuda_cape_multi / xcape_cuda_srh_multi: argument errors are reported before any device is touched, an
    empty grid is a no-op, and a device that does not exist is an error of that device, not a silent skip."""
    L = lib.lib()
    z = np.zeros(256, np.float32)
    zd = np.zeros(256, np.float64)
    zi = np.zeros(256, np.int32)
    p = lambda a: a.ctypes.data  # noqa: E731
    one = (C.c_int * 1)(0)

    def cape(ncol=8, nlev=1, source=1, devices=one, nd=1):
        return L.xcape_cuda_cape_multi(p(z), p(z), p(z), p(z), p(z), p(z), C.c_int64(ncol), nlev, 0, 0, 0, source, 1,
                                       C.c_float(500.), C.c_float(500.), None, p(z), p(z), p(zi), p(z), None, None, 0, devices, nd)

    def srh(ncol=8, devices=one, nd=1, precision=0):
        return L.xcape_cuda_srh_multi(*([p(z)] * 10), C.c_int64(ncol), 1, 0, 0, 0, C.c_double(3000.), C.c_double(2.), None,
                                      p(zd), p(zd), None, None, None, precision, devices, nd)
    assert cape(nd=0) == lib.ERR_ARG and b'devices' in L.xcape_cuda_last_error()
    assert cape(devices=None) == lib.ERR_ARG
    assert cape(source=7) == lib.ERR_ARG
    assert cape(ncol=0) == lib.OK
    assert srh(nd=0) == lib.ERR_ARG
    assert srh(precision=2) == lib.ERR_ARG
    assert srh(ncol=0) == lib.OK
    if lib.device_count() < 1:
        rc = cape(ncol=8)                                  # no GPU here: the shard's device cannot be selected
        assert rc != lib.OK and b'device 0' in L.xcape_cuda_last_error()


def test_missing_extension_fails_loudly(lib, monkeypatch):
    monkeypatch.setattr(lib, '_lib', None)
    monkeypatch.setattr(lib, 'LIB_PATH', '/nonexistent/libxcape_b200.so')
    with pytest.raises(ImportError):
        lib.lib()


def test_cuda_method_without_gpu_raises_not_falls_back(lib):
    """On a box without a CUDA device method='cuda' must raise — never compute on the CPU."""
    if lib.device_count() > 0:
        pytest.skip('a GPU is present')
    from xcape_b200 import core
    from xcape_b200.synthetic import make_soundings
    d = make_soundings('C1', cols=(0, 32))
    with pytest.raises(Exception) as ei:
        core.calc_cape(d['p'], d['t'], d['td'], d['ps'], d['ts'], d['tds'], vertical_lev='sigma', method='cuda')
    assert isinstance(ei.value, (lib.XcapeCudaError, ValueError))
