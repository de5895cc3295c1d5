"""Known-answer tests from the reference's own test-suite (test/test_krdem1.f90, test/test_krdem2.f90), run
through the ORACLE's DASKR (Krylov option) on the CPU: integration accuracy against the analytic solution and
root locations, with the reference tests' acceptance values.  This is the one place where the reference's own
tests pin the restated host logic (DASKR, ddstp, dnedk/dnsk, dslvk/dspigm on 1x1 and 2x2 systems, DRCHEK/DROOTS)."""
import ctypes as C

import numpy as np

from tests import krdem


def _oracle_call(L):
    L.o_daskr.argtypes = [C.c_void_p] * 21
    L.o_daskr.restype = None
    rpar, ipar = np.zeros(4), np.zeros(4, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    r = lambda x: C.cast(C.byref(x), C.c_void_p)  # noqa: E731
    fn = lambda f: C.cast(f, C.c_void_p) if f is not None else None  # noqa: E731

    def call(res, neq, t, y, yd, tout, info, rtol, atol, rwork, iwork, jac, psol, rt, nrt, jroot):
        neq_c, nrt_c, idid = C.c_int(neq), C.c_int(nrt), C.c_int(0)
        tc, toutc = C.c_double(t), C.c_double(tout)
        rt_a, at_a = np.atleast_1d(np.asarray(rtol, dtype=np.float64)), np.atleast_1d(np.asarray(atol, dtype=np.float64))
        lrw, liw = C.c_long(rwork.size), C.c_long(iwork.size)
        L.o_daskr(fn(res), r(neq_c), r(tc), p(y), p(yd), r(toutc), p(info), p(rt_a), p(at_a), r(idid), p(rwork), r(lrw),
                  p(iwork), r(liw), p(rpar), p(ipar), fn(jac), fn(psol), fn(rt), r(nrt_c), p(jroot))
        return tc.value, idid.value

    return call


def test_krdem1_oracle(oracle):
    ero, roots, iwork = krdem.run_krdem1(_oracle_call(oracle.L), krdem.HostArrays())
    print("krdem1: max error/atol %.2f, roots %s, NST %d NRE %d NRTE %d" % (
        ero, [(round(t, 6), j.tolist()) for t, j in roots], iwork[10], iwork[11], iwork[35]))
    krdem.check_krdem1(ero, roots)


def test_krdem2_oracle(oracle):
    roots, iwork, y = krdem.run_krdem2(_oracle_call(oracle.L), krdem.HostArrays())
    print("krdem2: roots %s, NST %d NRE %d NJE %d NRTE %d" % (roots, iwork[10], iwork[11], iwork[12], iwork[35]))
    krdem.check_krdem2(roots)
