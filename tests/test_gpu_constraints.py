"""Constraint checking in the initial-condition calculation (info(10) = 1, 3; DCNSTR / DCNST0,
src/daskr_new.f90:569-713, daskr.f:1477-1494, 1608-1622), GPU vs oracle on the food web.

ICNSTR = 2 (strictly positive) on every unknown.  The flat predator guess is far from the consistent value, so the
first Newton steps of the IC calculation change the predators by more than rlx = 40 %: DCNSTR shortens them.  Both
sides must take the same decisions (same counters) and a y that violates the constraints must be refused."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NP, MX, MY = 1, 20, 20
NS = 2 * NP
NEQ = NS * MX * MY
TOUTS = [1e-8, 1e-6, 1e-4, 1e-2]


def _info(info10):
    info = np.zeros(20, dtype=np.int32)
    info[[10, 11, 13, 14, 15]] = 1
    info[9] = info10
    return info


def run_gpu(dk, info10, pred_ic, icn=2):
    iwork = np.zeros(40 + 2 * NEQ, dtype=np.int32)
    lid = 40 + NEQ if info10 in (1, 3) else 40
    dk.setup_rbdpre(MX, MY, NS, NP, lid, iwork)
    if info10 in (1, 3):
        iwork[40:40 + NEQ] = icn
    dk.web_setpar(NP, MX, MY)
    info = _info(info10)
    rwork, rpar, ipar = np.zeros(60), np.zeros(1), np.array([1, 0], dtype=np.int32)
    c, cdot = dk.web_cinit(NEQ, pred_ic)
    t, idid = dk.daskr(dk.WEB_RES, NEQ, 0.0, c, cdot, TOUTS[0], info, 1e-5, 1e-5, rwork, iwork, rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)
    ic_counters = iwork[:40].copy()
    ys = []
    if idid >= 0:
        info[10] = 0
        for tout in TOUTS:
            t, idid = dk.daskr(dk.WEB_RES, NEQ, t, c, cdot, tout, info, 1e-5, 1e-5, rwork, iwork, rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)
            ys.append(c.copy())
    return idid, np.array(ys), ic_counters, iwork[:40].copy()


def run_oracle(oracle, info10, pred_ic, icn=2):
    L = oracle.L
    L.o_web_setpar(NP, MX, MY)
    iwork = np.zeros(40 + 3 * NEQ + 16, dtype=np.int32)
    lid = 40 + NEQ if info10 in (1, 3) else 40
    L.o_setup_rbdpre(MX, MY, NS, NP, lid, iwork.ctypes.data_as(C.c_void_p))
    if info10 in (1, 3):
        iwork[40:40 + NEQ] = icn
    rwork = np.zeros(104 + 19 * NEQ + int(iwork[26]) + 64)
    info = _info(info10)
    rpar, ipar = np.zeros(NEQ), np.array([1, 0], dtype=np.int32)
    c, cdot = oracle.web_cinit(NP, MX, MY, pred_ic=pred_ic)
    L.o_daskr.argtypes = [C.c_void_p] * 21
    L.o_daskr.restype = None
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    r = lambda x: C.cast(C.byref(x), C.c_void_p)  # noqa: E731
    fn = lambda f: C.cast(f, C.c_void_p)  # noqa: E731

    def call(t, tout):
        neq, nrt, idid, jroot = C.c_int(NEQ), C.c_int(0), C.c_int(0), C.c_int(0)
        tc, toutc, rt, at = C.c_double(t), C.c_double(tout), C.c_double(1e-5), C.c_double(1e-5)
        lrw, liw = C.c_long(rwork.size), C.c_long(iwork.size)
        L.o_daskr(fn(L.o_web_res), r(neq), r(tc), p(c), p(cdot), r(toutc), p(info), r(rt), r(at), r(idid), p(rwork), r(lrw),
                  p(iwork), r(liw), p(rpar), p(ipar), fn(L.o_web_jac), fn(L.o_web_psol), None, r(nrt), r(jroot))
        return tc.value, idid.value

    t, idid = call(0.0, TOUTS[0])
    ic_counters = iwork[:40].copy()
    ys = []
    if idid >= 0:
        info[10] = 0
        for tout in TOUTS:
            t, idid = call(t, tout)
            ys.append(c.copy())
    return idid, np.array(ys), ic_counters, iwork[:40].copy()


@pytest.mark.parametrize("info10,pred_ic", [(1, 1e5), (3, 1e5), (1, 3e5), (1, 2e4)])
def test_ic_with_constraints(dk, oracle, info10, pred_ic):
    g = run_gpu(dk, info10, pred_ic)
    o = run_oracle(oracle, info10, pred_ic)
    print("info(10)=%d pred=%g: idid %d/%d  IC counters NRE/NNI/NLI/NPS gpu %s cpu %s" % (
        info10, pred_ic, g[0], o[0], g[2][[11, 18, 19, 20]], o[2][[11, 18, 19, 20]]))
    assert g[0] == o[0]
    assert np.array_equal(g[2][[11, 12, 18, 19, 20]], o[2][[11, 12, 18, 19, 20]])        # the IC calculation itself
    if o[0] >= 0:
        w = 1.0 / (1e-5 * np.abs(o[1]) + 1e-5)
        assert np.sqrt(np.mean(((g[1] - o[1]) * w) ** 2, axis=1)).max() <= 10.0
        assert np.array_equal(g[3][[10, 12, 18]], o[3][[10, 12, 18]])


def test_constraints_change_the_ic_iteration(dk):
    """The constraint really acts: with it the IC Newton takes shortened steps (different counters than without)."""
    a = run_gpu(dk, 0, 3e5)
    b = run_gpu(dk, 1, 3e5)
    assert a[0] >= 0 and b[0] >= 0
    assert not np.array_equal(a[2][[11, 18, 19]], b[2][[11, 18, 19]])


def test_inconsistent_initial_y_is_refused(dk):
    with pytest.raises(dk.DaskrError, match="violates a constraint"):
        run_gpu(dk, 1, 1e5, icn=-2)       # "strictly negative" demanded of positive concentrations
    with pytest.raises(dk.DaskrError, match="ICNSTR"):
        run_gpu(dk, 1, 1e5, icn=5)
