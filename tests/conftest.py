import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def soundings():
    """Reference fixture dataset_soundings (test/fixtures.py:1165-3441): 13 columns x 108 levels,
    level 0 = surface.  Extracted by tests/make_golden.py."""
    z = np.load(os.path.join(GOLDEN, 'ref_soundings.npz'))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope='session')
def era5pl():
    """Reference fixture dataset_ERA5pressurelevel (test/fixtures.py:28-1162): 17 columns x 37
    pressure levels."""
    z = np.load(os.path.join(GOLDEN, 'ref_era5pl.npz'))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope='session')
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


def snd_cape_args(s):
    P, T, Td = s['pressure'], s['temperature'], s['dewpoint']
    return (P[:, 1:], T[:, 1:], Td[:, 1:], P[:, 0], T[:, 0], Td[:, 0])


def snd_srh_args(s):
    P, T, Td, U, V = s['pressure'], s['temperature'], s['dewpoint'], s['u_wind_ms'], s['v_wind_ms']
    return (P[:, 1:], T[:, 1:], Td[:, 1:], U[:, 1:], V[:, 1:], P[:, 0], T[:, 0], Td[:, 0], U[:, 0], V[:, 0])


def era_cape_args(e):
    return (e['level'], e['lev_t'], e['lev_td'], e['surf_p'], e['surf_t'], e['surf_td'])


def era_srh_args(e):
    return (e['level'], e['lev_t'], e['lev_td'], e['lev_u'], e['lev_v'], e['surf_p'], e['surf_t'], e['surf_td'],
            e['surf_u'], e['surf_v'])


# the reference's own tolerances (test/test_core.py:62-65, 219-220): assert_almost_equal decimals
def close_decimal(a, b, decimal):
    np.testing.assert_almost_equal(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), decimal)
