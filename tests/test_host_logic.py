"""Host-side logic that needs no GPU: argument validation and reshaping in xcape_b200.core
(mirrors reference test/test_core.py:39-58 with method='dummy'), array/layout plumbing,
column sharding, the synthetic generator."""
import numpy as np
import pytest

from xcape_b200 import _array as A
from xcape_b200 import _lib, core
from xcape_b200.sharding import column_blocks, rank_block
from xcape_b200.synthetic import CONFIGS, make_soundings


@pytest.mark.parametrize('source,n_returns', [('surface', 2), ('most-unstable', 4)])
@pytest.mark.parametrize('vertical_lev', ['sigma', 'pressure'])
def test_calc_cape_shape_3d_dummy(source, n_returns, vertical_lev):
    rng = np.random.default_rng(0)
    p, t, td = (rng.random((5, 10, 20)) for _ in range(3))
    ps, ts, tds = (rng.random((5, 10)) for _ in range(3))
    args = (p, t, td, ps, ts, tds) if vertical_lev == 'sigma' else (np.ones(20), t, td, ps, ts, tds)
    if vertical_lev == 'pressure':
        # the dummy backend asserts p.ndim == 2 (core.py:116) and gets p as (nlev, 1)
        pass
    result = core.calc_cape(*args, source=source, vertical_lev=vertical_lev, method='dummy')
    assert len(result) == n_returns
    for r in result:
        assert r.shape == p.shape[:-1]


def test_argument_errors():
    rng = np.random.default_rng(0)
    p, t, td = (rng.random((4, 6, 9)) for _ in range(3))
    ps, ts, tds = (rng.random((4, 6)) for _ in range(3))
    with pytest.raises(ValueError, match='Too few'):
        core.calc_cape(p, t, td, ps, ts, vertical_lev='sigma', method='dummy')
    with pytest.raises(ValueError, match='Too many'):
        core.calc_cape(p, t, td, ps, ts, tds, tds, vertical_lev='sigma', method='dummy')
    with pytest.raises(ValueError, match='vertical_lev'):
        core.calc_cape(p, t, td, ps, ts, tds, vertical_lev='eta', method='dummy')
    with pytest.raises(ValueError, match='P should be 1d'):
        core.calc_cape(p, t, td, ps, ts, tds, vertical_lev='pressure', method='dummy')
    with pytest.raises(ValueError, match='same shape'):
        core.calc_cape(p, t[:, :, :8], td, ps, ts, tds, vertical_lev='sigma', method='dummy')
    with pytest.raises(ValueError, match='same shape'):
        core.calc_cape(p, t, td, ps, ts[:3], tds, vertical_lev='sigma', method='dummy')
    with pytest.raises(ValueError, match='invalid method'):
        core.calc_cape(p, t, td, ps, ts, tds, vertical_lev='sigma', method='numba')
    with pytest.raises(KeyError):
        core.calc_cape(p, t, td, ps, ts, tds, vertical_lev='sigma', source='nope', method='dummy')
    with pytest.raises(KeyError):
        core.calc_cape(p, t, td, ps, ts, tds, vertical_lev='sigma', adiabat='nope', method='dummy')
    with pytest.raises(ValueError):
        core.calc_srh(p, t, td, vertical_lev='sigma')
    with pytest.raises(ImportError):
        core.calc_cape(p, t, td, ps, ts, tds, vertical_lev='sigma', method='fortran')


def test_columns_views_are_zero_copy_and_layouts():
    t = np.arange(4 * 6 * 9, dtype=np.float32).reshape(4, 6, 9)
    (t2,) = core._columns_2d([t], -1)
    assert t2.shape == (9, 24) and np.shares_memory(t2, t)
    assert A.layout_of_2d(t2) == _lib.LEVEL_LAST                 # reference layout: each column contiguous
    tm = np.ascontiguousarray(np.moveaxis(t, -1, 0))
    (t3,) = core._columns_2d([tm], 0)
    assert t3.shape == (9, 24) and np.shares_memory(t3, tm)
    assert A.layout_of_2d(t3) == _lib.LEVEL_MAJOR
    assert np.array_equal(t2, t3)
    assert A.layout_of_2d(t2[:, ::2]) is None
    f3, f1, p, dt, layout, mem, _ = A.prepare_fields([t2, t2[::1]], [np.zeros(24)], p=np.zeros(9, np.float32))
    assert dt == _lib.F64 and layout == _lib.LEVEL_LAST and mem == _lib.MEM_HOST     # mixed dtypes promote
    f3, *_rest = A.prepare_fields([t2[:, ::2], t2[:, 1::2]], [np.zeros(12, np.float32)])
    assert A.layout_of_2d(f3[0]) == _lib.LEVEL_LAST and _rest[2] == _lib.F32         # strided -> dense copy
    assert core._grid_shape((9,), -1) == (1,) and core._grid_shape((4, 6, 9), -1) == (4, 6)
    assert core._grid_shape((9, 4, 6), 0) == (4, 6)


def test_float32_fields_with_wider_1d_arguments_stay_float32_and_zero_copy():
    """a float64 / integer pressure axis or float64 surface values next to float32 3-D fields (the usual
    ERA5 case) must not drag the fields through a float64 host copy — f2py casts each argument separately."""
    t = np.random.default_rng(0).normal(size=(37, 1000)).astype(np.float32)      # level-major, dense
    td = t - 1
    for p in (np.arange(37.), np.arange(37), np.arange(37, dtype=np.float32)):
        for ps in (np.zeros(1000), np.zeros(1000, np.float32)):
            f3, f1, p_, dt, layout, mem, ref = A.prepare_fields([t, td], [ps, ps, ps], p=p, dtype_from='fields')
            assert dt == _lib.F32 and layout == _lib.LEVEL_MAJOR
            assert f3[0] is t and f3[1] is td                                   # zero-copy
            assert all(a.dtype == np.float32 for a in f1) and p_.dtype == np.float32
    # a float64 3-D field still promotes the group (the C ABI takes one dtype)
    *_, dt, _, _, _ = A.prepare_fields([t, td.astype(np.float64)], [np.zeros(1000, np.float32)], dtype_from='fields')
    assert dt == _lib.F64
    from xcape_b200.cape_cuda import _wider_than_fields
    assert _wider_than_fields([np.arange(37.), np.zeros(4, np.float32)], [t, td])
    assert not _wider_than_fields([np.arange(37, dtype=np.float32), np.zeros(4, np.float32)], [t, td])
    assert not _wider_than_fields([np.arange(37.)], [t.astype(np.float64)])


@pytest.mark.parametrize('ncol', [0, 1, 127, 128, 129, 1000, 721 * 1440, 24 * 721 * 1440])
@pytest.mark.parametrize('n', [1, 2, 3, 4, 8])
def test_column_blocks_partition(ncol, n):
    blocks = column_blocks(ncol, n)
    assert len(blocks) == n and blocks[0][0] == 0 and blocks[-1][1] == ncol
    for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [b - a for a, b in blocks]
    assert sum(sizes) == ncol and max(sizes) - min(sizes) <= 128 + 127
    assert all(a % 128 == 0 for a, _ in blocks if a < ncol)
    assert [rank_block(ncol, r, n) for r in range(n)] == blocks


def test_synthetic_is_deterministic_and_shardable():
    a = make_soundings('C2', cols=(3000, 9500))
    b = make_soundings('C2', cols=(0, 12000))
    for k in ('t', 'td', 'u', 'v', 'ps', 'ts', 'tds'):
        assert a[k].dtype == np.float32
        assert np.array_equal(a[k], b[k][3000:9500])
    assert a['p'].shape == (37,) and a['t'].shape == (6500, 37)
    c = make_soundings('C3', cols=(0, 100))
    assert c['p'].shape == (100, 50) and (np.diff(c['p'], axis=1) < 0).all()
    assert (a['tds'] <= 27.0).all() and (a['ts'] > 0).all() and (a['td'] <= a['t']).all()
    mix = make_soundings('C2', cols=(0, 20000), active=False)
    assert 0.2 < (mix['ts'] <= 0).mean() < 0.7
    s = make_soundings('C1', shuffle=True)
    u = make_soundings('C1')
    assert not np.array_equal(s['ps'], u['ps'])
    assert CONFIGS['C2']['grid'] == (721, 1440) and CONFIGS['C5']['nlev'] == 137
    d5 = make_soundings('C5', cols=(0, 8))
    assert d5['p'].shape == (8, 137) and d5['p'][:, -1].max() < 0.2


def test_level_order_detection():
    """core._top_first: 'auto' reads the storage order of the level axis off the pressure array."""
    from xcape_b200 import core
    up = np.array([1000., 850., 500., 100.])
    assert core._top_first(up, -1, 'auto') is False and core._top_first(up[::-1], -1, 'auto') is True
    p3 = np.broadcast_to(up, (3, 5, 4))
    assert core._top_first(p3, -1, 'auto') is False and core._top_first(p3[..., ::-1], -1, 'auto') is True
    pm = np.broadcast_to(up[:, None, None], (4, 3, 5))
    assert core._top_first(pm, 0, 'auto') is False and core._top_first(pm[::-1], 0, 'auto') is True
    assert core._top_first(up, -1, 'top_first') is True and core._top_first(up[::-1], -1, 'surface_first') is False
    with pytest.raises(ValueError):
        core._top_first(up, -1, 'upside-down')
    # the dummy backend ignores the order but accepts the keyword
    t = np.zeros((6, 4), np.float32)
    r = core.calc_cape(up[::-1].astype(np.float32), t, t, t[:, 0], t[:, 0], t[:, 0], vertical_lev='pressure',
                       method='dummy', level_order='auto')
    assert r[0].shape == (6,)


def test_stream_pipeline_order_errors_and_early_stop(tmp_path):
    """xcape_b200.stream: steps come back in order whatever the worker/reader counts, memory-mapped
    steps are read on the reader threads, loader exceptions reach the consumer, an abandoned
    generator winds its threads down."""
    import threading
    from xcape_b200.stream import _materialise, stream_cape, stream_srh
    steps = []
    for k in range(9):
        np.save(tmp_path / f't{k}.npy', np.full((6, 5, 4), k, np.float32))
        m = np.load(tmp_path / f't{k}.npy', mmap_mode='r')
        steps.append((m, m, m, m[..., 0], m[..., 0], m[..., 0]))
    got, borrowed = _materialise(steps[3])
    assert borrowed == []
    assert all(type(a) is np.ndarray and a.flags['C_CONTIGUOUS'] for a in got) and got[3].shape == (6, 5)
    before = threading.active_count()
    for devices, readers, prefetch in (([0], 1, 1), ([0, 1, 2], 2, 2), ([0, 1], 4, 3)):
        out = list(stream_cape(steps, devices=devices, readers=readers, prefetch=prefetch, method='dummy',
                               source='most-unstable'))
        assert len(out) == 9 and all(len(o) == 4 and o[0].shape == (6, 5) for o in out)
    from xcape_b200.stream import _stream
    import time

    def tag(*arrays, device, **kw):                # later steps finish first
        time.sleep(0.002 * (9 - int(arrays[0].ravel()[0])))
        return int(arrays[0].ravel()[0]), device
    out = list(_stream(tag, steps, [0, 1, 2], 3, 2, {}))
    assert [k for k, _ in out] == list(range(9)) and {d for _, d in out} <= {0, 1, 2}
    assert list(stream_cape([], method='dummy')) == []
    g = stream_cape([lambda s=s: s for s in steps], method='dummy')
    assert next(g)[0].shape == (6, 5)
    g.close()

    def bad():
        raise OSError('disk gone')
    with pytest.raises(OSError):
        list(stream_cape([steps[0], bad, steps[1]], method='dummy'))
    with pytest.raises(ValueError):
        stream_cape(steps, devices=[], method='dummy')
    with pytest.raises(ValueError):
        list(stream_srh([steps[0]], method='cuda'))          # arity error surfaces from calc_srh
    assert threading.active_count() == before


def test_cpulist_parsing_and_numa_binding_is_harmless_without_a_gpu():
    from xcape_b200.sharding import bind_host_to_gpu, parse_cpulist
    assert parse_cpulist('0-3,8,10-11\n') == {0, 1, 2, 3, 8, 10, 11}
    assert parse_cpulist('5') == {5} and parse_cpulist('') == set()
    import os
    before = os.sched_getaffinity(0)
    assert bind_host_to_gpu(0) is None            # no CUDA device here: nothing changes
    assert os.sched_getaffinity(0) == before


def test_gpu_cpu_affinity_from_nvidia_smi_topo():
    """bench.py / sharding bind a rank to the CPUs next to its GPU; where sysfs hides the NUMA node (containers)
    the set comes from `nvidia-smi topo -m`."""
    from xcape_b200.sharding import format_cpulist, gpu_cpu_affinity, parse_cpulist
    txt = ("\t\x1b[4mGPU0\tGPU1\tNIC0\tCPU Affinity\tNUMA Affinity\tGPU NUMA ID\x1b[0m\n"
           "GPU0\t X \tNV18\tSYS\t0-15,64-79\t0\t\tN/A\nGPU1\tNV18\t X \tPIX\t16-31\t1\t\tN/A\n"
           "NIC0\tSYS\tPIX\t X \t\t\t\t\n\nLegend:\n  X = Self\n")
    assert gpu_cpu_affinity(0, txt) == parse_cpulist('0-15,64-79')
    assert format_cpulist(gpu_cpu_affinity(1, txt)) == '16-31'
    assert gpu_cpu_affinity(2, txt) is None and gpu_cpu_affinity(0, 'garbage') is None
    assert format_cpulist({0, 1, 2, 5, 7, 8}) == '0-2,5,7-8'


def test_level_order_auto_skips_masked_columns_and_raises_when_undetermined():
    """'auto' used to read column 0 only — a NaN / fill value there made a top-first field run upside-down."""
    p = np.linspace(1000, 100, 10)
    assert core._top_first(p, -1, 'auto') is False and core._top_first(p[::-1], -1, 'auto') is True
    P = np.broadcast_to(p, (5, 7, 10)).copy()
    P[0, 0, :] = np.nan                               # first column masked
    P[0, 1, :] = 9.96921e36                           # second column a fill value
    assert core._top_first(P, -1, 'auto') is False and core._top_first(P[..., ::-1], -1, 'auto') is True
    assert core._top_first(np.moveaxis(P[..., ::-1], -1, 0), 0, 'auto') is True
    with pytest.raises(ValueError):
        core._top_first(np.full((3, 4), np.nan), -1, 'auto')
    with pytest.raises(ValueError):
        core._top_first(np.array([500., 500.]), -1, 'auto')
