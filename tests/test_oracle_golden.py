"""The oracle against every golden vector the reference's tests hold for the path
(reference test/test_core.py:68-286).  CPU only."""
import os

import numpy as np
import pytest

from conftest import close_decimal, era_cape_args, era_srh_args, snd_cape_args, snd_srh_args

TMODES = [0, 1, 2]  # LIBM, CR, SPEC


@pytest.mark.parametrize('tmode', TMODES)
@pytest.mark.parametrize('source,pinc,gc,gi', [
    ('surface', 100, 'SB_CAPE_pinc100', 'SB_CIN_pinc100'),
    ('mixed-layer', 1000, 'ML_CAPE_pinc1000_mldepth500', 'ML_CIN_pinc1000_mldepth500'),
    ('most-unstable', 100, 'MU_CAPE_pinc100', 'MU_CIN_pinc100')])
def test_cape_sigma_goldens(oracle_mod, soundings, tmode, source, pinc, gc, gi):
    # test_core.py:68-116
    r = oracle_mod.calc_cape_ref(*snd_cape_args(soundings), source=source, ml_depth=500., adiabat='pseudo-liquid',
                                 pinc=pinc, vertical_lev='sigma', tmode=tmode)
    close_decimal(r[0], soundings[gc], 0)
    close_decimal(r[1], soundings[gi], 0)
    # the survey's probe found these goldens are reproduced EXACTLY by a float32 restatement
    assert np.array_equal(r[0], soundings[gc].astype(np.float32))
    assert np.array_equal(r[1], soundings[gi].astype(np.float32))
    if source == 'most-unstable':
        assert np.array_equal(r[2], soundings['MU_lv_pinc100'].astype(np.int32))
        assert np.array_equal(r[3], soundings['MU_z_pinc100'].astype(np.float32))


@pytest.mark.parametrize('tmode', TMODES)
@pytest.mark.parametrize('source,gc,gi', [('surface', 'capesp500', 'cinsp500'),
                                          ('mixed-layer', 'capeml300p500', 'cinml300p500'),
                                          ('most-unstable', 'capemup500', 'cinmup500')])
def test_cape_pressure_goldens(oracle_mod, era5pl, tmode, source, gc, gi):
    # test_core.py:120-170
    r = oracle_mod.calc_cape_ref(*era_cape_args(era5pl), source=source, ml_depth=300, adiabat='pseudo-liquid',
                                 pinc=500, vertical_lev='pressure', tmode=tmode)
    close_decimal(r[0], era5pl['surf_' + gc], 0)
    close_decimal(r[1], era5pl['surf_' + gi], 0)
    # per-mode bounds (SURVEY §7.1 asked for <= 2e-3 with the libm the reference links): glibc arithmetic
    # reproduces the fixture to 1.1e-3 J/kg; correctly-rounded and SPEC transcendentals reproduce the
    # surface and mixed-layer goldens EXACTLY and differ on one most-unstable column by 0.0123 J/kg
    # (a one-ulp expf difference from glibc's, amplified by the parcel ascent)
    dc = np.abs(r[0] - era5pl['surf_' + gc]).max()
    di = np.abs(r[1] - era5pl['surf_' + gi]).max()
    if tmode == 0:
        assert dc < 2e-3 and di < 2e-3
    elif source == 'most-unstable':
        assert dc < 1.5e-2 and di == 0.0
    else:
        assert dc == 0.0 and di == 0.0


def test_pres_lev_pos_matches_fixture(oracle_mod, era5pl):
    plp = oracle_mod.pres_lev_pos(era5pl['surf_p'], era5pl['level'][:, None])
    assert set(np.unique(plp)) <= {2, 3}        # surface p ~ 968-979 hPa


def test_srh_sigma_goldens(oracle_mod, soundings):
    # test_core.py:173-229
    r = oracle_mod.calc_srh_ref(*snd_srh_args(soundings), depth=3000, vertical_lev='sigma', output_var='all')
    assert len(r) == 8
    close_decimal(r[0], soundings['SRH03_model_lev_rm'], 5)
    close_decimal(r[1], soundings['SRH03_model_lev_lm'], 5)


def test_srh_pressure_goldens(oracle_mod, era5pl):
    # test_core.py:231-286
    r = oracle_mod.calc_srh_ref(*era_srh_args(era5pl), depth=3000, vertical_lev='pressure', output_var='srh')
    close_decimal(r[0], era5pl['surf_srh_rm'], 5)
    close_decimal(r[1], era5pl['surf_srh_lm'], 5)


def test_stdheight_golden(oracle_mod, soundings):
    # AGLH_model_lev is carried by the fixture (not asserted by the reference) — pins stdheight
    a = snd_cape_args(soundings)
    p2, t2, td2 = oracle_mod._to2d(a[0], a[1], a[2])
    H, Hs = oracle_mod.stdheight(p2, t2, td2, a[3], a[4], a[5], 0, 1, 2., 1)
    ref = soundings['AGLH_model_lev']
    assert np.nanmax(np.abs(H.T - ref[:, 1:])) < 1e-9
    assert np.array_equal(Hs, ref[:, 0])


def L_one(oracle_mod):
    return float(oracle_mod.lib().xcape_ref_logf(1.0, 2))


def test_spec_math_is_a_valid_libm(oracle_mod):
    """SPEC exp/log/pow agree with the correctly-rounded binary32 result at least as often as
    glibc's own functions do.  log and pow have binary64 cores (mismatch rate ~2^-20); exp is the binary32
    float-float version (1.0e-4 of calls, never more than one ulp; glibc's expf: 6e-4)."""
    rng = np.random.default_rng(0)

    def ulps(a, b):
        return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64))

    for lo, hi, fn in ((-53.6, 7.1, 'exp'), (-87.0, 88.0, 'exp'), (-0.05, 0.05, 'exp'), (-0.2, 0.2, 'exp_small'),
                       (-0.015625, 0.015625, 'exp_small'), (-1e-3, 1e-3, 'exp_small')):
        x = rng.uniform(lo, hi, 1_000_000).astype(np.float32)
        cr = oracle_mod.vec_math('exp', x, tmode=1)
        sp = oracle_mod.vec_math(fn, x, tmode=2)
        gl = oracle_mod.vec_math('exp', x, tmode=0)
        assert (sp != cr).mean() < 2.5e-4, (lo, hi, fn)
        assert ulps(sp, cr).max() <= 1
        assert (sp != cr).sum() <= (gl != cr).sum() + 20, 'SPEC exp must not be further from CR than glibc expf'
    # subnormal results and the overflow edge: one rounding only
    x = np.concatenate([rng.uniform(-104.5, -87.0, 200_000), rng.uniform(88.0, 89.0, 50_000)]).astype(np.float32)
    cr = oracle_mod.vec_math('exp', x, tmode=1)
    sp = oracle_mod.vec_math('exp', x, tmode=2)
    assert np.array_equal(np.isinf(sp), np.isinf(cr))
    fin = np.isfinite(cr)
    assert ulps(sp[fin], cr[fin]).max() <= 1
    x = np.exp(rng.uniform(-8, 8, 500_000)).astype(np.float32)
    assert (oracle_mod.vec_math('log', x, tmode=2) != oracle_mod.vec_math('log', x, tmode=1)).mean() < 1e-4
    # the table-driven log near its cancellation-prone spot: EVERY binary32 within 2^-10 of 1, the interval
    # edges of the table, and pressure ratios p2/p1 of the ascent
    lo, hi = np.float32(1 - 2.0 ** -10).view(np.uint32), np.float32(1 + 2.0 ** -10).view(np.uint32)
    x = np.arange(lo, hi + 1, dtype=np.uint32).view(np.float32)
    assert (oracle_mod.vec_math('log', x, tmode=2) != oracle_mod.vec_math('log', x, tmode=1)).sum() == 0
    edges = np.float32(0.6875) + np.arange(0, 176, dtype=np.float32) * np.float32(2.0 ** -8)     # 0.6875 .. 1.375
    x = np.concatenate([np.nextafter(edges, np.float32(0)), edges, np.nextafter(edges, np.float32(2))]).astype(np.float32)
    x = np.concatenate([x * np.float32(2.0 ** k) for k in (-126, -20, -1, 0, 1, 30, 126)])
    assert (oracle_mod.vec_math('log', x, tmode=2) != oracle_mod.vec_math('log', x, tmode=1)).sum() == 0
    x = rng.uniform(0.9, 1.0, 500_000).astype(np.float32)
    assert (oracle_mod.vec_math('log', x, tmode=2) != oracle_mod.vec_math('log', x, tmode=1)).mean() < 1e-4
    assert L_one(oracle_mod) == 0.0
    x = rng.uniform(0.0005, 1.1, 500_000).astype(np.float32)
    y = np.float32(287.04) / np.float32(1005.7)
    assert (oracle_mod.vec_math('pow', x, y, tmode=2) != oracle_mod.vec_math('pow', x, y, tmode=1)).mean() < 1e-4
    L = oracle_mod.lib()
    assert L.xcape_ref_expf(-800.0, 2) == 0.0 and np.isinf(L.xcape_ref_expf(800.0, 2))
    assert np.isnan(L.xcape_ref_expf(float('nan'), 2)) and np.isnan(L.xcape_ref_logf(-1.0, 2))
    assert L.xcape_ref_logf(0.0, 2) == -np.inf


def test_skip_and_status_counters(oracle_mod, soundings):
    """ts <= 0 degC gate (CAPE_CODE_model_lev.f90:77): sounding #13 has Ts = -4.9 degC."""
    (c, ci, lv, z), cnt = oracle_mod.calc_cape_ref(*snd_cape_args(soundings), source='most-unstable', pinc=100,
                                                   vertical_lev='sigma', counters=True)
    assert soundings['temperature'][12, 0] < 0
    assert c[12] == 0 and ci[12] == 0 and lv[12] == 0 and z[12] == 0
    assert cnt['status'][12] == 1 and cnt['n_iter'][12] == 0
    assert (cnt['status'][:12] == 0).all() and (cnt['n_iter'][:12] > 100).all()


def test_contracted_oracle_is_a_small_perturbation(oracle_mod, soundings):
    a = oracle_mod.calc_cape_ref(*snd_cape_args(soundings), source='surface', pinc=100, vertical_lev='sigma')
    b = oracle_mod.calc_cape_ref(*snd_cape_args(soundings), source='surface', pinc=100, vertical_lev='sigma',
                                 contract=True)
    assert np.abs(a[0] - b[0]).max() < 1.0


def test_dewpoint_from_q_formula_against_the_era5_fixture():
    """The reference ships no q -> Td routine, but its ERA5 fixture carries q next to td (unused by its tests).
    The formula behind xcape_cuda_dewpoint_from_q (inverse of getqvs) reproduces the fixture's td wherever the
    air is above freezing — to 0.014 K on levels, 0.008 K at the surface; below 0 degC the fixture was produced
    with an ice / mixed-phase saturation law (differences 0.6-4.3 K), which is why the op stays 'parity unpinned'."""
    import oracle
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_era5pl.npz'))
    td_s = oracle.dewpoint_from_q_ref(z['surf_p'], 1e-3 * z['surf_q'], q_min=0.0)          # fixture q is in g/kg
    assert np.abs(td_s - z['surf_td']).max() < 1e-2
    td_l = oracle.dewpoint_from_q_ref(z['level'][None, :].astype(np.float64), 1e-3 * z['lev_q'], q_min=0.0)
    d = np.abs(td_l - z['lev_td'])
    warm = z['lev_t'] >= 0.0
    assert warm.sum() > 200 and d[warm].max() < 2e-2
    assert d[~warm].min() > 0.5 and d.max() < 4.5
