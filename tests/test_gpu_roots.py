"""Rootfinding (DRCHEK / DROOTS, src/daskr.f:2494-2910), GPU vs oracle.

The food-web example as shipped runs with nrt = 1, R = average(c1) - 20 (example_web.f90:600-617);
the heat example with nrt = 2, R = max(u) - {0.1, 0.01} (example_heat.f90:91-109).  Both sides must
report the same roots (same count, same jroot, root times equal to a tight relative tolerance -- the
root function is a reduction, so only its summation order differs) and continue on the same step
sequence afterwards.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_daskr_loop(L, res, jac, psol, rt, nrt, neq, y, yd, touts, info, rtol, atol, rwork, iwork, rpar, ipar, ic):
    L.o_daskr.argtypes = [C.c_void_p] * 21
    L.o_daskr.restype = None
    jroot = np.zeros(max(nrt, 1), dtype=np.int32)
    fn = lambda f: C.cast(f, C.c_void_p)  # noqa: E731
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    r = lambda x: C.cast(C.byref(x), C.c_void_p)  # noqa: E731

    def call(t, tout):
        neq_c, nrt_c, idid = C.c_int(neq), C.c_int(nrt), C.c_int(0)
        tc, toutc = C.c_double(t), C.c_double(tout)
        rt_c, at_c = C.c_double(rtol), C.c_double(atol)
        lrw, liw = C.c_long(rwork.size), C.c_long(iwork.size)
        L.o_daskr(fn(res), r(neq_c), r(tc), p(y), p(yd), r(toutc), p(info), r(rt_c), r(at_c), r(idid), p(rwork), r(lrw),
                  p(iwork), r(liw), p(rpar), p(ipar), fn(jac), fn(psol), fn(rt), r(nrt_c), p(jroot))
        return tc.value, idid.value

    t = 0.0
    if ic:
        t, idid = call(t, touts[0])
        assert idid == 4
        info[10] = 0
    ys, cnts, found = [], [], []
    for tout in touts:
        while True:
            t, idid = call(t, tout)
            assert idid >= 0
            if idid == 5:
                found.append((t, jroot.copy(), iwork[:40].copy()))
                continue
            break
        ys.append(y.copy())
        cnts.append(iwork[:40].copy())
    return dict(y=np.array(ys), counters=np.array(cnts), roots=found)


def oracle_web_roots(oracle, np_, mx, my, touts):
    L = oracle.L
    ns = 2 * np_
    neq = ns * mx * my
    L.o_web_setpar(np_, mx, my)
    iwork = np.zeros(40 + 2 * neq + 16, dtype=np.int32)
    L.o_setup_rbdpre(mx, my, ns, np_, 40, iwork.ctypes.data_as(C.c_void_p))
    rwork = np.zeros(104 + 19 * neq + int(iwork[26]) + 64 + 3)
    info = np.zeros(20, dtype=np.int32)
    info[[10, 11, 13, 14, 15]] = 1
    rpar, ipar = np.zeros(neq), np.array([1, 0], dtype=np.int32)
    c, cdot = oracle.web_cinit(np_, mx, my, pred_ic=1e5 * np_)   # the product's solve_web default
    return _oracle_daskr_loop(L, L.o_web_res, L.o_web_jac, L.o_web_psol, L.o_web_rt, 1, neq, c, cdot, touts, info,
                              1e-5, 1e-5, rwork, iwork, rpar, ipar, ic=True)


def oracle_heat_roots(oracle, m, touts):
    L = oracle.L
    neq = (m + 2) ** 2
    L.o_heat_setpar(m)
    iwork = np.zeros(40 + 2 * neq + 16, dtype=np.int32)
    L.o_setup_rbdpre(m + 2, m + 2, 1, 1, 0, iwork.ctypes.data_as(C.c_void_p))
    rwork = np.zeros(104 + 19 * neq + int(iwork[26]) + 64 + 6)
    info = np.zeros(20, dtype=np.int32)
    info[[11, 14]] = 1
    rpar, ipar = np.zeros(2), np.array([0, 0, neq, m], dtype=np.int32)
    u, ud = np.zeros(neq), np.zeros(neq)
    L.o_heat_uinit.argtypes = [C.c_void_p, C.c_void_p]
    L.o_heat_uinit(u.ctypes.data_as(C.c_void_p), ud.ctypes.data_as(C.c_void_p))
    return _oracle_daskr_loop(L, L.o_heat_res, L.o_heat_jac, L.o_heat_psol, L.o_heat_rt, 2, neq, u, ud, touts, info,
                              0.0, 1e-5, rwork, iwork, rpar, ipar, ic=False)


def _same_roots(g, o, ttol):
    assert len(g["roots"]) == len(o["roots"]), (g["roots"], o["roots"])
    for (tg, jg, cg), (to, jo, co) in zip(g["roots"], o["roots"]):
        assert abs(tg - to) <= ttol * abs(to), (tg, to)
        assert np.array_equal(jg[: jo.size], jo)
        assert cg[10] == co[10]                           # NST at the root
        # NRTE: DROOTS iterates until the bracket is narrower than 100*uround*(|tn|+|h|); its last few
        # iterations see the root function at rounding level, where the summation order of the reduction
        # decides the sign, so the iteration count may differ by a few evaluations
        assert abs(int(cg[35]) - int(co[35])) <= 16


def test_web_root_as_shipped(dk, oracle):
    """example_web.f90 as shipped: 20x20 mesh, np=1, nrt=1.  One root (c1 average rising through 20)."""
    touts = [1e-8, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1]
    g = dk.solve_web(1, 20, 20, touts, roots=True)
    o = oracle_web_roots(oracle, 1, 20, 20, touts)
    assert len(o["roots"]) == 1 and o["roots"][0][1][0] == 1
    _same_roots(g, o, 1e-9)
    w = 1.0 / (1e-5 * np.abs(o["y"]) + 1e-5)
    assert np.sqrt(np.mean(((g["y"] - o["y"]) * w) ** 2, axis=1)).max() <= 10.0
    sel = [10, 11, 12, 13, 14, 15, 18, 20]               # NST NRE NJE NETF NCFN NCFL NNI NPS
    assert np.array_equal(g["counters"][:7][:, sel], o["counters"][:7][:, sel])
    # the root value itself: average(c1) == 20 at the reported time (interpolated solution returned at idid=5)
    g0 = dk.solve_web(1, 20, 20, touts, roots=False)
    assert np.array_equal(g0["counters"][:, 10], g["counters"][:, 10])   # root checks do not change the step sequence


def test_web_root_np2_larger_mesh(dk, oracle):
    touts = [1e-6, 1e-4, 1e-2, 5e-2]
    g = dk.solve_web(2, 32, 32, touts, roots=True)
    o = oracle_web_roots(oracle, 2, 32, 32, touts)
    _same_roots(g, o, 1e-8)
    w = 1.0 / (1e-5 * np.abs(o["y"]) + 1e-5)
    assert np.sqrt(np.mean(((g["y"] - o["y"]) * w) ** 2, axis=1)).max() <= 10.0


def test_heat_roots(dk, oracle):
    """example_heat.f90's two root functions with the Krylov/rbdpre pairing: max(u) falls through 0.1 then 0.01."""
    m = 30
    touts = [0.01 * 2 ** k for k in range(1, 8)]
    g = dk.solve_heat(m, touts, roots=True)
    o = oracle_heat_roots(oracle, m, touts)
    assert len(o["roots"]) >= 1
    # the two solutions agree to ~1e-2 x atol (1e-7 in u); max(u) decays at 2*pi^2*max(u) ~ 0.2/s, so the
    # crossing times agree to ~1e-6 relative
    _same_roots(g, o, 1e-4)
    assert np.abs(g["y"] - o["y"]).max() <= 1e-4
