"""Checkpoint-style continuation: export the device-side state + the caller's rwork / iwork,
tear the library down, bring it up again, import, continue -- the continued run must take exactly the steps of
the uninterrupted one (bitwise equal outputs and counters)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NP, MX, MY = 1, 24, 20
NS = 2 * NP
TOUTS = [1e-6, 1e-4, 1e-3, 1e-2, 1e-1]


def _setup(dk):
    probe = np.zeros(64, dtype=np.int32)
    dk.setup_rbdpre(MX, MY, NS, NP, 0, probe)
    neq = NS * MX * MY
    iwork = np.zeros(40 + neq, dtype=np.int32)
    dk.setup_rbdpre(MX, MY, NS, NP, 40, iwork)
    dk.web_setpar(NP, MX, MY)
    return neq, iwork


def _run(dk, jpre, break_after=None, info16=1, vector_tol=False):
    neq, iwork = _setup(dk)
    info = np.zeros(20, dtype=np.int32)
    info[[10, 11, 13, 14]] = 1
    info[15] = info16                                 # info(16): error test on the differential variables only (VT) or all (VT = WT)
    rwork, rpar, ipar = np.zeros(60), np.zeros(1), np.array([jpre, 0], dtype=np.int32)
    c, cdot = dk.web_cinit(neq)
    rtol = atol = 1e-5
    if vector_tol:
        info[1] = 1                                   # info(2) = 1: RTOL / ATOL are vectors
        rtol, atol = np.full(neq, 1e-5), np.full(neq, 1e-5)
    args = lambda: (rwork, iwork, rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)  # noqa: E731
    t, idid = dk.daskr(dk.WEB_RES, neq, 0.0, c, cdot, TOUTS[0], info, rtol, atol, *args())
    assert idid == 4
    info[10] = 0
    ys, cnts = [], []
    for k, tout in enumerate(TOUTS):
        if break_after is not None and k == break_after:
            # ---- checkpoint: device state + the caller's arrays ----
            saved = dict(state=dk.state_export(), rwork=rwork.copy(), iwork=iwork.copy(), info=info.copy(), t=t,
                         c=c.copy(), cdot=cdot.copy())
            dk.finalize()
            dk.init(0)
            neq, _ = _setup(dk)                       # a fresh process would do exactly this
            dk.state_import(saved["state"])
            rwork, iwork, info = saved["rwork"], saved["iwork"], saved["info"]
            c, cdot, t = saved["c"], saved["cdot"], saved["t"]
        t, idid = dk.daskr(dk.WEB_RES, neq, t, c, cdot, tout, info, rtol, atol, *args())
        assert idid == 3
        ys.append(c.copy())
        cnts.append(iwork[:40].copy())
    return np.array(ys), np.array(cnts)


@pytest.mark.parametrize("jpre", [1, 3])
def test_export_import_continues_bitwise(dk, jpre):
    y0, c0 = _run(dk, jpre)
    y1, c1 = _run(dk, jpre, break_after=2)
    assert np.array_equal(c0, c1)
    bad = np.argwhere(y0 != y1)
    assert bad.size == 0, "first differences (output, unknown): %s, max |dy| = %.3e" % (bad[:5].tolist(), np.max(np.abs(y0 - y1)))


@pytest.mark.parametrize("vector_tol", [False, True])
def test_export_import_with_vt_aliased_to_wt(dk, vector_tol):
    """info(16) = 0: the reference's VT is the same rwork segment as WT (src/daskr.f:1416-1427).  After an import the
    alias must hold again (it used to be set on the initial-call path only: every error norm was then taken against
    vt = 0, the error test always passed and the step doubled every step)."""
    y0, c0 = _run(dk, 1, info16=0, vector_tol=vector_tol)
    y1, c1 = _run(dk, 1, break_after=2, info16=0, vector_tol=vector_tol)
    assert np.array_equal(c0, c1)
    assert np.array_equal(y0, y1)
    ya, ca = _run(dk, 1, info16=1)
    assert not np.array_equal(c0, ca) or not np.array_equal(y0, ya)   # the two error tests really differ on this problem


def test_tolerances_are_reread_on_continuation(dk, oracle):
    """The reference reads RTOL / ATOL in DDAWTS at every step, so a caller may change vector tolerances between
    continuation calls (and must, after idid = -2).  GPU and oracle, same change at the same output time."""
    import ctypes as C

    neq, iwork = _setup(dk)
    info = np.zeros(20, dtype=np.int32)
    info[[1, 10, 11, 13, 14, 15]] = 1
    rwork, rpar, ipar = np.zeros(60), np.zeros(1), np.array([1, 0], dtype=np.int32)
    c, cdot = dk.web_cinit(neq)
    rtol, atol = np.full(neq, 1e-5), np.full(neq, 1e-5)
    t, idid = dk.daskr(dk.WEB_RES, neq, 0.0, c, cdot, TOUTS[0], info, rtol, atol, rwork, iwork, rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)
    assert idid == 4
    info[10] = 0
    for k, tout in enumerate(TOUTS):
        if k == 2:
            rtol[:] = 1e-8
            atol[:] = 1e-8
        t, idid = dk.daskr(dk.WEB_RES, neq, t, c, cdot, tout, info, rtol, atol, rwork, iwork, rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)
        assert idid == 3
    g_cnt, g_y = iwork[:40].copy(), c.copy()

    L = oracle.L
    L.o_web_setpar(NP, MX, MY)
    iwo = np.zeros(40 + 2 * neq + 16, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    L.o_setup_rbdpre(MX, MY, NS, NP, 40, p(iwo))
    lrw = 104 + 19 * neq + int(iwo[26]) + 64
    rwo, rpo = np.zeros(lrw), np.zeros(neq)
    co, cdo = oracle.web_cinit(NP, MX, MY)
    info = np.zeros(20, dtype=np.int32)
    info[[1, 10, 11, 13, 14, 15]] = 1
    rtol, atol = np.full(neq, 1e-5), np.full(neq, 1e-5)
    L.o_daskr.argtypes = [C.c_void_p] * 21
    L.o_daskr.restype = None

    def call(t, tout):
        neqc, nrt, idid, jroot = C.c_int(neq), C.c_int(0), C.c_int(0), C.c_int(0)
        tc, toutc = C.c_double(t), C.c_double(tout)
        lrwc, liwc = C.c_long(lrw), C.c_long(iwo.size)
        r = lambda x: C.cast(C.byref(x), C.c_void_p)  # noqa: E731
        L.o_daskr(C.cast(L.o_web_res, C.c_void_p), r(neqc), r(tc), p(co), p(cdo), r(toutc), p(info), p(rtol), p(atol), r(idid),
                  p(rwo), r(lrwc), p(iwo), r(liwc), p(rpo), p(ipar), C.cast(L.o_web_jac, C.c_void_p),
                  C.cast(L.o_web_psol, C.c_void_p), None, r(nrt), r(jroot))
        return tc.value, idid.value

    t, idid = call(0.0, TOUTS[0])
    assert idid == 4
    info[10] = 0
    for k, tout in enumerate(TOUTS):
        if k == 2:
            rtol[:] = 1e-8
            atol[:] = 1e-8
        t, idid = call(t, tout)
        assert idid == 3
    idx = [10, 11, 12, 13, 14, 15, 18, 19, 20]
    print("counters gpu %s cpu %s" % (g_cnt[idx].tolist(), iwo[idx].tolist()))
    assert np.array_equal(g_cnt[idx], iwo[idx])
    w = 1.0 / (1e-8 * np.abs(co) + 1e-8)
    assert np.sqrt(np.mean(((g_y - co) * w) ** 2)) <= 10.0
    assert g_cnt[10] > 60                             # the tighter tolerance really took effect (the loose run takes ~40 steps)


def test_async_output_equals_blocking_output(dk):
    """DKB_OPT_ASYNC_OUT: y / yprime of every intermediate-output return arrive through a snapshot and a second stream;
    after dkb_output_wait() they are the bits the blocking transfer delivers."""
    outs = {}
    for mode in (0, 1):
        neq, iwork = _setup(dk)
        dk.set_option(dk.OPT_ASYNC_OUT, mode)
        info = np.zeros(20, dtype=np.int32)
        info[[2, 10, 11, 13, 14, 15]] = 1            # info(3) = 1: return after every step
        rwork, rpar, ipar = np.zeros(60), np.zeros(1), np.array([3, 0], dtype=np.int32)
        c, cdot = dk.web_cinit(neq)
        t, idid = dk.daskr(dk.WEB_RES, neq, 0.0, c, cdot, 1e-3, info, 1e-5, 1e-5, rwork, iwork, rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)
        info[10] = 0
        ys = []
        for _ in range(25):
            t, idid = dk.daskr(dk.WEB_RES, neq, t, c, cdot, 1e-3, info, 1e-5, 1e-5, rwork, iwork, rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)
            assert idid in (1, 3)
            dk.output_wait()
            ys.append((t, c.copy(), cdot.copy()))
            if idid == 3:
                break
        outs[mode] = ys
        dk.set_option(dk.OPT_ASYNC_OUT, 0)
    assert len(outs[0]) == len(outs[1]) > 5
    for a, b in zip(outs[0], outs[1]):
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_import_rejects_garbage(dk):
    _setup(dk)
    with pytest.raises(dk.DaskrError):
        dk.state_import(np.zeros(4096, dtype=np.uint8))
