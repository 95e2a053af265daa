"""ctypes front end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; ``xcape_b200`` never does.

The wrappers mimic what f2py does at the reference's L1 boundary:
inputs are cast to the routine's dtype and to Fortran ``(nk, n2)`` order (each column
contiguous), outputs are freshly allocated, zero-filled numpy arrays.

``cape(...)``, ``stdheight(...)`` and ``srh(...)`` have the signatures of the
reference's L2 shims (``cape_fortran.py:3``, ``stdheight.py:5``, ``srh.py:4``) so that
``calc_cape_ref`` / ``calc_srh_ref`` below restate ``core.py:261-332`` / ``473-542``.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBM, CR, SPEC = 0, 1, 2
_libs = {}


def build(force=False):
    """Compile the oracle shared objects with gcc (no GPU needed)."""
    tgt = [os.path.join(_HERE, 'libxcape_oracle.so'), os.path.join(_HERE, 'libxcape_oracle_fma.so')]
    src = os.path.join(_HERE, 'xcape_oracle.cpp')
    if force or any((not os.path.exists(t)) or os.path.getmtime(t) < os.path.getmtime(src) for t in tgt):
        subprocess.check_call(['make', '-C', _HERE, '-s', '-B', 'all'])


def lib(contract=False):
    key = bool(contract)
    if key not in _libs:
        name = 'libxcape_oracle_fma.so' if contract else 'libxcape_oracle.so'
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        assert L.xcape_ref_fp_contract() == int(key)
        L.xcape_ref_expf.restype = C.c_float
        L.xcape_ref_expf.argtypes = [C.c_float, C.c_int]
        L.xcape_ref_logf.restype = C.c_float
        L.xcape_ref_logf.argtypes = [C.c_float, C.c_int]
        L.xcape_ref_powf.restype = C.c_float
        L.xcape_ref_powf.argtypes = [C.c_float, C.c_float, C.c_int]
        _libs[key] = L
    return _libs[key]


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f(a, dt):
    """f2py-style coercion: dtype cast + Fortran-contiguous copy only when needed."""
    return np.asfortranarray(a, dtype=dt)


def vec_math(fn, x, y=None, tmode=SPEC):
    x = np.ascontiguousarray(x, np.float32)
    out = np.empty_like(x)
    L = lib()
    if fn == 'exp':
        L.xcape_ref_expf_v(_p(x), _p(out), C.c_int64(x.size), C.c_int(tmode))
    elif fn == 'exp_small':
        L.xcape_ref_expf_small_v(_p(x), _p(out), C.c_int64(x.size), C.c_int(tmode))
    elif fn == 'log':
        L.xcape_ref_logf_v(_p(x), _p(out), C.c_int64(x.size), C.c_int(tmode))
    else:
        y = np.ascontiguousarray(np.broadcast_to(y, x.shape), np.float32)
        L.xcape_ref_powf_v(_p(x), _p(y), _p(out), C.c_int64(x.size), C.c_int(tmode))
    return out


# --------------------------------------------------------------------------------------
# L1: the f2py routines
# --------------------------------------------------------------------------------------
def loopcape_ml(p3d, t3d, td3d, ps, ts, tds, pinc, source, ml_depth, adiabat,
                tmode=CR, nthreads=1, contract=False, counters=False):
    """CAPE_CODE_model_lev.pyf:6-24."""
    p3d, t3d, td3d = (_f(a, np.float32) for a in (p3d, t3d, td3d))
    ps, ts, tds = (np.ascontiguousarray(a, np.float32) for a in (ps, ts, tds))
    nk, n2 = t3d.shape
    assert p3d.shape == (nk, n2) and td3d.shape == (nk, n2) and ps.shape == (n2,)
    cape = np.zeros(n2, np.float32); cin = np.zeros(n2, np.float32)
    mulvl = np.zeros(n2, np.int32); zout = np.zeros(n2, np.float32)
    cnt = [np.zeros(n2, np.int32) for _ in range(3)] if counters else [None] * 3
    cond = np.full(n2, np.inf, np.float32) if counters else None
    lib(contract).xcape_ref_set_cond_buffer(_p(cond))
    rc = lib(contract).xcape_ref_loopcape_ml(
        _p(p3d), _p(t3d), _p(td3d), _p(ps), _p(ts), _p(tds), C.c_float(pinc), C.c_int(source),
        C.c_float(ml_depth), C.c_int(adiabat), C.c_int(nk), C.c_int64(n2), _p(cape), _p(cin),
        _p(mulvl), _p(zout), C.c_int(tmode), C.c_int(nthreads), _p(cnt[0]), _p(cnt[1]), _p(cnt[2]))
    lib(contract).xcape_ref_set_cond_buffer(None)
    if rc:
        raise ValueError('oracle: bad source/adiabat/pinc')
    if counters:
        return cape, cin, mulvl, zout, dict(n_iter=cnt[0], n_sub=cnt[1], status=cnt[2], cond_b=cond)
    return cape, cin, mulvl, zout


def loopcape_pl1d(t3d, td3d, p, ps, ts, tds, pinc, source, ml_depth, adiabat, start_3d,
                  tmode=CR, nthreads=1, contract=False, counters=False):
    """CAPE_CODE_pressure_lev.pyf:26-45."""
    t3d, td3d = (_f(a, np.float32) for a in (t3d, td3d))
    p = np.ascontiguousarray(np.asarray(p, np.float32).reshape(-1))
    ps, ts, tds = (np.ascontiguousarray(a, np.float32) for a in (ps, ts, tds))
    nk, n2 = t3d.shape
    assert p.shape == (nk,)
    start = np.ascontiguousarray(np.broadcast_to(start_3d, (n2,)), np.int32)
    cape = np.zeros(n2, np.float32); cin = np.zeros(n2, np.float32)
    mulvl = np.zeros(n2, np.int32); zout = np.zeros(n2, np.float32)
    cnt = [np.zeros(n2, np.int32) for _ in range(3)] if counters else [None] * 3
    cond = np.full(n2, np.inf, np.float32) if counters else None
    lib(contract).xcape_ref_set_cond_buffer(_p(cond))
    rc = lib(contract).xcape_ref_loopcape_pl1d(
        _p(t3d), _p(td3d), _p(p), _p(ps), _p(ts), _p(tds), C.c_float(pinc), C.c_int(source),
        C.c_float(ml_depth), C.c_int(adiabat), _p(start), C.c_int(nk), C.c_int64(n2), _p(cape),
        _p(cin), _p(mulvl), _p(zout), C.c_int(tmode), C.c_int(nthreads), _p(cnt[0]), _p(cnt[1]),
        _p(cnt[2]))
    lib(contract).xcape_ref_set_cond_buffer(None)
    if rc:
        raise ValueError('oracle: bad source/adiabat/pinc')
    if counters:
        return cape, cin, mulvl, zout, dict(n_iter=cnt[0], n_sub=cnt[1], status=cnt[2], cond_b=cond)
    return cape, cin, mulvl, zout


def loop_stdheight_ml(p, t, td, ps, ts, tds, hin, nthreads=1):
    """stdheight_2D_model_lev.pyf:6-19."""
    p, t, td = (_f(a, np.float64) for a in (p, t, td))
    ps, ts, tds, hin = (np.ascontiguousarray(a, np.float64) for a in (ps, ts, tds, hin))
    nk, nx = t.shape
    H = np.zeros((nk, nx), np.float64, order='F'); Hs = np.zeros(nx, np.float64)
    lib().xcape_ref_loop_stdheight_ml(_p(p), _p(t), _p(td), _p(ps), _p(ts), _p(tds), _p(hin),
                                      C.c_int(nk), C.c_int64(nx), _p(H), _p(Hs), C.c_int(nthreads))
    return H, Hs


def loop_stdheight_pl1d(t, td, p, ps, ts, tds, hin, start_3d, nthreads=1):
    """stdheight_2D_pressure_lev.pyf:21-35."""
    t, td = (_f(a, np.float64) for a in (t, td))
    p = np.ascontiguousarray(np.asarray(p, np.float64).reshape(-1))
    ps, ts, tds, hin = (np.ascontiguousarray(a, np.float64) for a in (ps, ts, tds, hin))
    nk, nx = t.shape
    start = np.ascontiguousarray(np.broadcast_to(start_3d, (nx,)), np.float64)
    H = np.zeros((nk, nx), np.float64, order='F'); Hs = np.zeros(nx, np.float64)
    lib().xcape_ref_loop_stdheight_pl1d(_p(t), _p(td), _p(p), _p(ps), _p(ts), _p(tds), _p(hin),
                                        _p(start), C.c_int(nk), C.c_int64(nx), _p(H), _p(Hs),
                                        C.c_int(nthreads))
    return H, Hs


def bunkers_loop(u, v, aglh, us, vs, aglhs, start_3d=None, nthreads=1):
    """Bunkers_model_lev.pyf:8-21 (start_3d None) / Bunkers_pressure_lev.pyf:6-20."""
    u, v, aglh = (_f(a, np.float32) for a in (u, v, aglh))
    us, vs, aglhs = (np.ascontiguousarray(a, np.float32) for a in (us, vs, aglhs))
    nk, n2 = u.shape
    start = None if start_3d is None else np.ascontiguousarray(np.broadcast_to(start_3d, (n2,)), np.float32)
    RM = np.zeros((2, n2), np.float32, order='F'); LM = np.zeros((2, n2), np.float32, order='F')
    M6 = np.zeros((2, n2), np.float32, order='F')
    lib().xcape_ref_bunkers_loop(_p(u), _p(v), _p(aglh), _p(us), _p(vs), _p(aglhs), _p(start),
                                 C.c_int(nk), C.c_int64(n2), _p(RM), _p(LM), _p(M6), C.c_int(nthreads))
    return RM, LM, M6


def loop_sreh(u, v, aglh, us, vs, aglhs, cu_rm, cv_rm, cu_lm, cv_lm, top, start_3d=None, nthreads=1):
    """SREH_model_lev.pyf:6-23 (start_3d None) / SREH_pressure_lev.pyf:6-24."""
    u, v, aglh = (_f(a, np.float64) for a in (u, v, aglh))
    us, vs, aglhs, cu_rm, cv_rm, cu_lm, cv_lm = (
        np.ascontiguousarray(a, np.float64) for a in (us, vs, aglhs, cu_rm, cv_rm, cu_lm, cv_lm))
    nk, n2 = u.shape
    start = None if start_3d is None else np.ascontiguousarray(np.broadcast_to(start_3d, (n2,)), np.float64)
    srm = np.zeros(n2, np.float64); slm = np.zeros(n2, np.float64)
    lib().xcape_ref_loop_sreh(_p(u), _p(v), _p(aglh), _p(us), _p(vs), _p(aglhs), _p(cu_rm), _p(cv_rm),
                              _p(cu_lm), _p(cv_lm), C.c_double(top), _p(start), C.c_int(nk),
                              C.c_int64(n2), _p(srm), _p(slm), C.c_int(nthreads))
    return srm, slm


# --------------------------------------------------------------------------------------
# L2: the per-operator shims (cape_fortran.py:3-69, stdheight.py:5-38, srh.py:4-66)
# --------------------------------------------------------------------------------------
def cape(p_2d, t_2d, td_2d, p_s, t_s, td_s, flag_1d, pres_lev_pos, source, ml_depth, adiabat,
         pinc, type_grid, **kw):
    if type_grid == 1:
        return loopcape_ml(p_2d, t_2d, td_2d, p_s, t_s, td_s, pinc, source, ml_depth, adiabat, **kw)
    if type_grid == 2 and flag_1d == 1:
        return loopcape_pl1d(t_2d, td_2d, p_2d, p_s, t_s, td_s, pinc, source, ml_depth, adiabat,
                             pres_lev_pos, **kw)
    raise ValueError('type_grid/flag_1d')


def stdheight(p_2d, t_2d, td_2d, p_s, t_s, td_s, flag_1d, pres_lev_pos, aglh0, type_grid, nthreads=1):
    nlev, ngrid = t_2d.shape
    aglh_in = np.ones(ngrid) * aglh0 if np.isscalar(aglh0) else aglh0
    if type_grid == 1:
        return loop_stdheight_ml(p_2d, t_2d, td_2d, p_s, t_s, td_s, aglh_in, nthreads)
    return loop_stdheight_pl1d(t_2d, td_2d, p_2d, p_s, t_s, td_s, aglh_in, pres_lev_pos, nthreads)


def srh(u_2d, v_2d, aglh_2d, u_s, v_s, aglh_s, pres_lev_pos, depth, type_grid, output, nthreads=1):
    start = None if type_grid == 1 else pres_lev_pos
    rm, lm, m6 = bunkers_loop(u_2d, v_2d, aglh_2d, u_s, v_s, aglh_s, start, nthreads)
    srm, slm = loop_sreh(u_2d, v_2d, aglh_2d, u_s, v_s, aglh_s, rm[0, :], rm[1, :], lm[0, :], lm[1, :],
                         depth, start, nthreads)
    if output == 1:
        return srm, slm
    return srm, slm, rm, lm, m6


# --------------------------------------------------------------------------------------
# L3: restatement of core._calc_cape_numpy / _calc_srh_numpy (core.py:261-332, 473-542)
# --------------------------------------------------------------------------------------
_SOURCE = {'surface': 1, 'most-unstable': 2, 'mixed-layer': 3}
_ADIABAT = {'pseudo-liquid': 1, 'reversible-liquid': 2, 'pseudo-ice': 3, 'reversible-ice': 4}
_VLEV = {'sigma': 1, 'pressure': 2}


def _to2d(*a):
    a2 = [np.atleast_2d(x) for x in a]
    sh = a2[0].shape
    n = int(np.prod(sh[:-1]))
    return [np.reshape(x, (n, sh[-1])).transpose() for x in a2]


def pres_lev_pos(p_s1d, p_2d):
    """core.py:286-289 verbatim semantics (numpy masked argmin)."""
    temp_index = (p_s1d - p_2d)
    return np.ma.masked_less(temp_index, 0).argmin(axis=0) + 1


def calc_cape_ref(p, t, td, ps, ts, tds, source='surface', ml_depth=500., adiabat='pseudo-liquid',
                  pinc=500., vertical_lev='sigma', **kw):
    p, t, td, ps, ts, tds = (np.asarray(a) for a in (p, t, td, ps, ts, tds))
    shape = t.shape
    p_s, t_s, td_s = (np.reshape(np.atleast_1d(a), (-1,)) for a in (ps, ts, tds))
    if p.ndim == 1:
        t_2d, td_2d = _to2d(t, td)
        p_2d = _to2d(p)[0]
        flag_1d = 1
        plp = pres_lev_pos(p_s, p_2d)
    else:
        p_2d, t_2d, td_2d = _to2d(p, t, td)
        flag_1d = 0
        plp = 1
    out = cape(p_2d, t_2d, td_2d, p_s, t_s, td_s, flag_1d, plp, _SOURCE[source], ml_depth,
               _ADIABAT[adiabat], pinc, _VLEV[vertical_lev], **kw)
    tgt = (1,) if len(shape) == 1 else shape[:-1]
    res = [np.reshape(a, tgt) for a in out[:4]]
    if kw.get('counters'):
        return res, out[4]
    return res if _SOURCE[source] == 2 else res[:2]


def calc_srh_ref(p, t, td, u, v, ps, ts, tds, us, vs, depth=3000, vertical_lev='sigma',
                 output_var='srh', nthreads=1):
    p, t, td, u, v, ps, ts, tds, us, vs = (np.asarray(a) for a in (p, t, td, u, v, ps, ts, tds, us, vs))
    shape = t.shape
    p_s, t_s, td_s, u_s, v_s = (np.reshape(np.atleast_1d(a), (-1,)) for a in (ps, ts, tds, us, vs))
    if p.ndim == 1:
        t_2d, td_2d, u_2d, v_2d = _to2d(t, td, u, v)
        p_2d = _to2d(p)[0]
        flag_1d = 1
        plp = pres_lev_pos(p_s, p_2d)
    else:
        p_2d, t_2d, td_2d, u_2d, v_2d = _to2d(p, t, td, u, v)
        flag_1d = 0
        plp = 1
    tg = _VLEV[vertical_lev]
    aglh_2d, aglh_s = stdheight(p_2d, t_2d, td_2d, p_s, t_s, td_s, flag_1d, plp, 2., tg, nthreads)
    out = srh(u_2d, v_2d, aglh_2d, u_s, v_s, aglh_s, plp, depth, tg, 1 if output_var == 'srh' else 2, nthreads)
    tgt = (1,) if len(shape) == 1 else shape[:-1]
    res = [np.reshape(a, tgt) for a in out[:2]]
    if output_var == 'srh':
        return res
    tgt2 = (2,) if len(shape) == 1 else (2,) + shape[:-1]
    # reference quirk kept: a plain C-order reshape of the (2, ncol) array (core.py:65-79)
    rm, lm, m6 = (np.reshape(a, tgt2) for a in out[2:])
    return res + [rm[0], rm[1], lm[0], lm[1], m6[0], m6[1]]


def dewpoint_from_q_ref(p_hpa, q, q_min=1e-10):
    """float64 numpy statement of xcape_cuda_dewpoint_from_q (parity unpinned: the reference has no
    such routine).  Inverse of getqvs (CAPE_CODE_model_lev.f90:570-581): r = q/(1-q), e = p r/(eps+r),
    L = ln(e/6.112), Td = 243.5 L/(17.67 - L) degC."""
    p_hpa = np.asarray(p_hpa, np.float64)
    q = np.maximum(np.asarray(q, np.float64), q_min) if q_min > 0 else np.asarray(q, np.float64)
    r = q / (1.0 - q)
    e = p_hpa * r / (287.04 / 461.5 + r)
    with np.errstate(divide='ignore', invalid='ignore'):
        L = np.log(e / 6.112)
        return 243.5 * L / (17.67 - L)


def qvs_ref(p_pa, t_k):
    """getqvs (CAPE_CODE_model_lev.f90:570-581) in float64: saturation mixing ratio over liquid."""
    es = 611.2 * np.exp(17.67 * (t_k - 273.15) / (t_k - 29.65))
    return (287.04 / 461.5) * es / (p_pa - es)
