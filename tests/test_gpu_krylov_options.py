"""Branches of the Krylov path that the default settings never reach, GPU (through the C ABI) against the oracle.

  * dkb_dslvk called directly -- the entry a Fortran dnsk binds (src/daskr_new.f90:2982) -- with user callbacks on
    device arrays, on a random nonsymmetric LINEAR system, for MAXL / KMP / NRMAX settings that include incomplete
    orthogonalisation (KMP < MAXL, dspigm:3329-3350 and the `dl` restart residual :3376-3402), MAXL = 10 and
    NRMAX = 0 (no restarts);
  * dorth's conditional reorthogonalisation (:3575-3588), forced by a system whose Krylov space has dimension 2,
    with the branch counters of both sides checked;
  * info(13) = 1 with iwork(24:26) / rwork(10) set by the user, whole food-web runs.
"""
import ctypes as C

import numpy as np
import pytest

from tests import krdem

pytestmark = pytest.mark.gpu

RES_T = C.CFUNCTYPE(None, *([C.c_void_p] * 8))
PSOL_T = C.CFUNCTYPE(None, *([C.c_void_p] * 15))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _linear_callbacks(m, mem, pre=None):
    """res: delta = M y.  psol: b = pre @ b (pre = None: no preconditioning)."""
    n = m.shape[0]

    def res(t, y, yd, cj, delta, ires, rpar, ipar):
        mem.put(delta, m @ mem.get(y, n))

    def psol(neq, t, y, yd, savr, wk, cj, wght, wp, iwp, b, eps, ier, rpar, ipar):
        if pre is not None:
            mem.put(b, pre @ mem.get(b, n))
        C.cast(ier, C.POINTER(C.c_int))[0] = 0

    return RES_T(res), PSOL_T(psol)


def _oracle_dslvk(oracle, m, pre, y, b, ewt, maxl, kmp, nrmax, eplin):
    n = m.shape[0]
    L = oracle.L
    L.o_dslvk.argtypes = [C.c_long, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double,
                          C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    L.o_dslvk.restype = None
    res_c, psol_c = _linear_callbacks(m, krdem.HostArrays(), pre)
    yd = np.zeros(n)
    savr = m @ y
    x = b.copy()
    w = ewt.copy()
    rwm = np.zeros(n * (maxl + 4) + (maxl + 1) * maxl + 2 * maxl + 64)
    iwm = np.zeros(64, dtype=np.int32)
    iwm[23], iwm[24], iwm[25] = maxl, kmp, nrmax
    iwm[28], iwm[29] = rwm.size - 7, 41
    ires, iersl, rhok = C.c_int(0), C.c_int(0), C.c_double(0)
    rpar, ipar = np.zeros(2), np.zeros(2, dtype=np.int32)
    sqrtn = np.sqrt(n)
    L.o_dslvk(n, _p(y), 0.0, _p(yd), _p(savr), _p(x), _p(w), _p(rwm), _p(iwm), C.cast(res_c, C.c_void_p), C.byref(ires),
              C.cast(psol_c, C.c_void_p), C.byref(iersl), 1.0, eplin, sqrtn, 1.0 / sqrtn, C.byref(rhok), _p(rpar), _p(ipar))
    return x, iwm[[11, 15, 19, 20]].copy(), ires.value, iersl.value, rhok.value


def _gpu_dslvk(dk, m, pre, y, b, ewt, maxl, kmp, nrmax, eplin):
    from daskr_b200.api import DeviceBuffer

    n = m.shape[0]
    dk.finalize()
    dk.init(0)                      # plain user problem: no mesh, no built-in preconditioner
    L = dk.lib()
    assert L.dkb_alloc_workspace(n, maxl, kmp, 0, 0) == 0, L.dkb_last_error()
    res_c, psol_c = _linear_callbacks(m, krdem.DeviceArrays(L), pre)
    d_y, d_yd = DeviceBuffer.from_host(y), DeviceBuffer.from_host(np.zeros(n))
    d_savr, d_x, d_w = DeviceBuffer.from_host(m @ y), DeviceBuffer.from_host(b), DeviceBuffer.from_host(ewt)
    iwm = np.zeros(64, dtype=np.int32)
    iwm[23], iwm[24], iwm[25] = maxl, kmp, nrmax
    neq, ires, iersl = C.c_int(n), C.c_int(0), C.c_int(0)
    t, cj, eps, rhok = C.c_double(0.0), C.c_double(1.0), C.c_double(eplin), C.c_double(0.0)
    sqrtn, rsqrtn = C.c_double(np.sqrt(n)), C.c_double(1.0 / np.sqrt(n))
    rpar, ipar = np.zeros(2), np.zeros(2, dtype=np.int32)
    L.dkb_dslvk.restype = None
    L.dkb_dslvk.argtypes = [C.c_void_p] * 19
    r = C.byref
    L.dkb_dslvk(r(neq), d_y.ptr, r(t), d_yd.ptr, d_savr.ptr, d_x.ptr, d_w.ptr, _p(iwm), C.cast(res_c, C.c_void_p), r(ires),
                C.cast(psol_c, C.c_void_p), r(iersl), r(cj), r(eps), r(sqrtn), r(rsqrtn), r(rhok), _p(rpar), _p(ipar))
    x = d_x.to_host(np.float64, (n,))
    w_after = d_w.to_host(np.float64, (n,))
    assert np.array_equal(w_after, ewt)                 # weights restored on exit (dslvk:3163)
    return x, iwm[[11, 15, 19, 20]].copy(), ires.value, iersl.value, rhok.value


@pytest.mark.parametrize("maxl,kmp,nrmax,eplin", [(5, 5, 5, 0.05), (5, 5, 5, 1e-6), (5, 2, 5, 1e-4), (6, 3, 2, 1e-3),
                                                  (10, 10, 0, 1e-6), (3, 3, 0, 0.05), (8, 8, 1, 1e-8)])
def test_dslvk_direct_against_oracle(dk, oracle, maxl, kmp, nrmax, eplin):
    n = 40
    rng = np.random.default_rng(12 + maxl * 7 + kmp)
    m = np.eye(n) + 0.35 * rng.standard_normal((n, n)) / np.sqrt(n)
    pre = np.diag(1.0 / np.diag(m))                      # Jacobi preconditioner (left), exercised through psol
    y, b = rng.standard_normal(n), rng.standard_normal(n)
    ewt = 1.0 / (0.05 * np.abs(y) + 0.05)
    xo, co, ireso, iersl_o, rhoko = _oracle_dslvk(oracle, m, pre, y, b, ewt, maxl, kmp, nrmax, eplin)
    xg, cg, iresg, iersl_g, rhokg = _gpu_dslvk(dk, m, pre, y, b, ewt, maxl, kmp, nrmax, eplin)
    print("maxl %d kmp %d nrmax %d eplin %g: counters [nre ncfl nli nps] gpu %s cpu %s, iersl %d/%d, rhok %.3e/%.3e"
          % (maxl, kmp, nrmax, eplin, cg, co, iersl_g, iersl_o, rhokg, rhoko))
    assert (iresg, iersl_g) == (ireso, iersl_o)
    assert np.array_equal(cg, co)
    assert np.allclose(xg, xo, rtol=1e-9, atol=1e-12 * np.abs(xo).max())
    assert abs(rhokg - rhoko) <= 1e-6 * max(abs(rhoko), 1e-300) + 1e-14
    if iersl_o == 0:                                     # converged: the true preconditioned residual meets eplin
        resid = (pre @ (m @ xg - b)) * ewt
        assert np.sqrt(np.mean(resid ** 2)) <= eplin * 1.01


def test_dorth_reorthogonalisation_branch(dk, oracle):
    """M = I + u w^T: the Krylov space of b has dimension 2, so at ll = 2 the new vector is rounding noise after
    Gram-Schmidt and `vnorm + snormw/1000 == vnorm` holds (dorth:3575).  With w orthogonal to the second basis
    vector, hes(1,2) is itself at rounding level, so the correction statements :3581-3585 run as well."""
    n = 40
    rng = np.random.default_rng(5)
    b = rng.standard_normal(n)
    v1 = b / np.linalg.norm(b)
    u = rng.standard_normal(n)
    up = u - (u @ v1) * v1
    w = rng.standard_normal(n)
    w = w - (w @ up) / (up @ up) * up                    # w . v2 = 0 (v2 is parallel to up)
    w = w / np.linalg.norm(w)
    m = np.eye(n) + np.outer(u, w) / np.linalg.norm(u)
    y = np.zeros(n)                                      # G(0) = 0: the difference quotient of datv is exactly M v
    ewt = np.full(n, 20.0)
    L = oracle.L
    L.o_reorth_stats.argtypes = [C.c_void_p]
    before = np.zeros(2, dtype=np.int64)
    L.o_reorth_stats(_p(before))
    xo, co, _, iersl_o, rhoko = _oracle_dslvk(oracle, m, None, y, b, ewt, 5, 5, 5, 1e-10)
    after = np.zeros(2, dtype=np.int64)
    L.o_reorth_stats(_p(after))
    xg, cg, _, iersl_g, rhokg = _gpu_dslvk(dk, m, None, y, b, ewt, 5, 5, 5, 1e-10)
    ge, gc = C.c_int64(0), C.c_int64(0)
    assert dk.lib().dkb_reorth_stats(C.byref(ge), C.byref(gc)) == 0
    print("reorth entered / corrections: cpu %s gpu (%d, %d); counters gpu %s cpu %s; iersl %d/%d"
          % ((after - before).tolist(), ge.value, gc.value, cg, co, iersl_g, iersl_o))
    assert (after - before)[0] >= 1 and ge.value >= 1    # both sides took the branch
    assert iersl_g == iersl_o
    assert np.array_equal(cg, co)
    xs = np.linalg.solve(m, b)
    assert np.allclose(xg, xs, rtol=1e-8, atol=1e-10) and np.allclose(xo, xs, rtol=1e-8, atol=1e-10)


def _web_run(side, dk_or_oracle, np_, mx, my, touts, maxl, kmp, nrmax, epli, jpre=1):
    """The food-web example with info(13) = 1 (user MAXL, KMP, NRMAX, EPLI); both sides through their DASKR entry."""
    ns = 2 * np_
    neq = ns * mx * my
    info = np.zeros(20, dtype=np.int32)
    info[[10, 11, 12, 13, 14, 15]] = 1
    ipar = np.array([jpre, 0], dtype=np.int32)
    if side == "gpu":
        dk = dk_or_oracle
        iwork = np.zeros(40 + neq, dtype=np.int32)
        dk.setup_rbdpre(mx, my, ns, np_, 40, iwork)
        dk.web_setpar(np_, mx, my)
        rwork, rpar = np.zeros(60), np.zeros(1)
        c, cdot = dk.web_cinit(neq, 1e5 * np_)

        def call(t, tout):
            return dk.daskr(dk.WEB_RES, neq, t, c, cdot, tout, info, 1e-5, 1e-5, rwork, iwork, rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)
    else:
        oracle = dk_or_oracle
        L = oracle.L
        L.o_web_setpar(np_, mx, my)
        iwork = np.zeros(40 + 2 * neq + 16, dtype=np.int32)
        L.o_setup_rbdpre(mx, my, ns, np_, 40, _p(iwork))
        lrw = 104 + (maxl + 24) * neq + int(iwork[26]) + (maxl + 3) * maxl + 64
        rwork, rpar = np.zeros(lrw), np.zeros(neq)
        c, cdot = oracle.web_cinit(np_, mx, my, 1e5 * np_)
        L.o_daskr.argtypes = [C.c_void_p] * 21
        L.o_daskr.restype = None

        def call(t, tout):
            neqc, nrt, idid, jroot = C.c_int(neq), C.c_int(0), C.c_int(0), C.c_int(0)
            tc, toutc = C.c_double(t), C.c_double(tout)
            rt, at = C.c_double(1e-5), C.c_double(1e-5)
            lrwc, liwc = C.c_long(lrw), C.c_long(iwork.size)
            r = lambda x: C.cast(C.byref(x), C.c_void_p)  # noqa: E731
            L.o_daskr(C.cast(L.o_web_res, C.c_void_p), r(neqc), r(tc), _p(c), _p(cdot), r(toutc), _p(info), r(rt), r(at), r(idid),
                      _p(rwork), r(lrwc), _p(iwork), r(liwc), _p(rpar), _p(ipar), C.cast(L.o_web_jac, C.c_void_p),
                      C.cast(L.o_web_psol, C.c_void_p), None, r(nrt), r(jroot))
            return tc.value, idid.value
    iwork[23], iwork[24], iwork[25] = maxl, kmp, nrmax
    rwork[9] = epli
    t, idid = call(0.0, touts[0])
    ys = []
    if idid >= 0:
        info[10] = 0
        for tout in touts:
            t, idid = call(t, tout)
            ys.append(c.copy())
            if idid < 0:
                break
    return idid, np.array(ys), iwork[:40].copy()


@pytest.mark.parametrize("maxl,kmp,nrmax,epli", [(5, 2, 5, 0.05), (10, 10, 5, 0.05), (4, 4, 0, 0.05), (8, 3, 2, 0.01), (1, 1, 5, 0.05)])
@pytest.mark.parametrize("np_,mx,my,jpre", [(1, 20, 20, 1), (2, 32, 16, 3)])
def test_info13_krylov_parameters(dk, oracle, maxl, kmp, nrmax, epli, np_, mx, my, jpre):
    touts = [1e-8, 1e-6, 1e-4, 1e-3]
    g = _web_run("gpu", dk, np_, mx, my, touts, maxl, kmp, nrmax, epli, jpre)
    o = _web_run("cpu", oracle, np_, mx, my, touts, maxl, kmp, nrmax, epli, jpre)
    idx = [10, 11, 12, 13, 14, 15, 18, 19, 20]
    print("maxl %d kmp %d nrmax %d epli %g np %d %dx%d jpre %d: idid %d/%d counters gpu %s cpu %s"
          % (maxl, kmp, nrmax, epli, np_, mx, my, jpre, g[0], o[0], g[2][idx].tolist(), o[2][idx].tolist()))
    assert g[0] == o[0]
    assert g[2][23] == maxl and g[2][24] == kmp and g[2][25] == nrmax
    n = min(len(g[1]), len(o[1]))
    assert n == len(o[1])
    if n:
        w = 1.0 / (1e-5 * np.abs(o[1][:n]) + 1e-5)
        err = np.sqrt(np.mean(((g[1][:n] - o[1][:n]) * w) ** 2, axis=1))
        assert err.max() <= 10.0, err
    # identical control flow until rounding-level differences flip a convergence test: the step counts agree closely
    assert abs(int(g[2][10]) - int(o[2][10])) <= max(2, 0.05 * o[2][10])
    assert abs(int(g[2][19]) - int(o[2][19])) <= max(3, 0.08 * o[2][19])
