"""The reference's own two test problems (test/test_krdem1.f90, test/test_krdem2.f90) set up for the Krylov option.

The reference runs them with the direct solver; the acceptance values of its tests are properties of the
integration and of the rootfinder, not of the linear solver:
  krdem1: |y - exp(-t^2+5t-4)| < 1e3*atol (atol = 1e-6) at every return; roots of r1 = y' at t = 2.5 and of
          r2 = ln y - 2.2491 at t = 2.47, 2.53, each located to 1e-3 (test_krdem1.f90:149-192);
  krdem2: roots of r = y1 (Van der Pol, mu = 100) at 81.17237787055 + (k-1)*81.41853556212, located to 1.0
          (test_krdem2.f90:205-217).
Here info(12) = 1: krdem1 with no preconditioner (psol = identity), krdem2 with the exact 2x2 iteration matrix
cj*I - dF/dy as preconditioner (jac stores it in wp, psol solves with it).  `Backend` hides whether the
length-NEQ arrays the callbacks get are host (oracle) or device (GPU library) memory.
"""
import ctypes as C
import math

import numpy as np

RES_T = C.CFUNCTYPE(None, *([C.c_void_p] * 8))
JACK_T = C.CFUNCTYPE(None, *([C.c_void_p] * 16))
PSOL_T = C.CFUNCTYPE(None, *([C.c_void_p] * 15))
RT_T = C.CFUNCTYPE(None, *([C.c_void_p] * 8))


def _dbl(addr, n=1):
    return np.ctypeslib.as_array(C.cast(addr, C.POINTER(C.c_double)), shape=(n,))


def _int(addr):
    return C.cast(addr, C.POINTER(C.c_int))


class HostArrays:
    """oracle: callback arrays are host memory"""

    def get(self, addr, n):
        return _dbl(addr, n).copy()

    def put(self, addr, vals):
        _dbl(addr, len(vals))[:] = vals


class DeviceArrays:
    """GPU library: callback arrays are device memory; move the 1-2 doubles through the C ABI's copy helpers"""

    def __init__(self, lib):
        self.L = lib

    def get(self, addr, n):
        out = np.empty(n)
        assert self.L.dkb_memcpy_d2h(out.ctypes.data_as(C.c_void_p), C.c_void_p(addr), 8 * n) == 0
        return out

    def put(self, addr, vals):
        a = np.ascontiguousarray(vals, dtype=np.float64)
        assert self.L.dkb_memcpy_h2d(C.c_void_p(addr), a.ctypes.data_as(C.c_void_p), a.nbytes) == 0


def krdem1_callbacks(mem):
    def res(t, y, yd, cj, delta, ires, rpar, ipar):
        tt = _dbl(t)[0]
        yv, ydv = mem.get(y, 1)[0], mem.get(yd, 1)[0]
        if yv <= 0.0:
            _int(ires)[0] = -1
            return
        mem.put(delta, [ydv - ((2 * math.log(yv) + 8.0) / tt - 5.0) * yv])

    def psol(neq, t, y, yd, savr, wk, cj, wght, wp, iwp, b, eps, ier, rpar, ipar):
        _int(ier)[0] = 0          # no preconditioning

    def rt(neq, t, y, yd, nrt, rval, rpar, ipar):
        yv, ydv = mem.get(y, 1)[0], mem.get(yd, 1)[0]
        _dbl(rval, 2)[:] = [ydv, math.log(yv) - 2.2491]

    return RES_T(res), None, PSOL_T(psol), RT_T(rt)


def krdem2_callbacks(mem):
    def res(t, y, yd, cj, delta, ires, rpar, ipar):
        yv, ydv = mem.get(y, 2), mem.get(yd, 2)
        mem.put(delta, [ydv[0] - yv[1], ydv[1] - (100 * (1.0 - yv[0] ** 2) * yv[1] - yv[0])])

    def jac(resf, ires, neq, t, y, yd, rewt, savr, wk, h, cj, wp, iwp, ier, rpar, ipar):
        yv, c = mem.get(y, 2), _dbl(cj)[0]
        # cj*I - dF/dy, row-major (test_krdem2.f90:45-66)
        mem.put(wp, [c - 0.0, -1.0, -(-200 * yv[0] * yv[1] - 1.0), c - 100 * (1.0 - yv[0] ** 2)])
        _int(ier)[0] = 0

    def psol(neq, t, y, yd, savr, wk, cj, wght, wp, iwp, b, eps, ier, rpar, ipar):
        a, rhs = mem.get(wp, 4).reshape(2, 2), mem.get(b, 2)
        mem.put(b, np.linalg.solve(a, rhs))
        _int(ier)[0] = 0

    def rt(neq, t, y, yd, nrt, rval, rpar, ipar):
        _dbl(rval, 1)[0] = mem.get(y, 1)[0]

    return RES_T(res), JACK_T(jac), PSOL_T(psol), RT_T(rt)


def run_krdem1(daskr_call, mem):
    """daskr_call(res, neq, t, y, yd, tout, info, rtol, atol, rwork, iwork, jac, psol, rt, nrt, jroot) -> (t, idid).
    Returns (max error / atol, [(t_root, jroot)], iwork)."""
    res, jac, psol, rt = krdem1_callbacks(mem)
    info = np.zeros(20, dtype=np.int32)
    info[11] = 1                       # info(12): Krylov
    y, yd = np.array([1.0]), np.array([3.0])
    rwork, iwork = np.zeros(400), np.zeros(100, dtype=np.int32)
    jroot = np.zeros(2, dtype=np.int32)
    t, tout, ero, roots = 1.0, 2.0, 0.0, []
    for _ in range(5):
        while True:
            t, idid = daskr_call(res, 1, t, y, yd, tout, info, 0.0, 1e-6, rwork, iwork, jac, psol, rt, 2, jroot)
            assert idid >= 0, idid
            ero = max(ero, abs(y[0] - math.exp(-t * t + 5 * t - 4.0)) / 1e-6)
            if idid != 5:
                tout += 1.0
                break
            roots.append((t, jroot.copy()))
    return ero, roots, iwork


def run_krdem2(daskr_call, mem):
    res, jac, psol, rt = krdem2_callbacks(mem)
    info = np.zeros(20, dtype=np.int32)
    info[1] = 1                        # info(2): vector tolerances
    info[11] = 1                       # Krylov
    info[14] = 1                       # info(15): jac supplied
    rtol, atol = np.array([1e-6, 1e-6]), np.array([1e-6, 1e-4])
    y, yd = np.array([2.0, 0.0]), np.array([0.0, -2.0])
    rwork, iwork = np.zeros(400), np.zeros(100, dtype=np.int32)
    iwork[26], iwork[27] = 4, 0        # iwork(27:28): lenwp, leniwp
    jroot = np.zeros(1, dtype=np.int32)
    t, tout, roots = 0.0, 20.0, []
    for _ in range(10):
        while True:
            t, idid = daskr_call(res, 2, t, y, yd, tout, info, rtol, atol, rwork, iwork, jac, psol, rt, 1, jroot)
            assert idid >= 0, idid
            if idid != 5:
                tout += 20.0
                break
            roots.append((t, int(jroot[0])))
    return roots, iwork, y.copy()


def check_krdem1(ero, roots):
    assert ero < 1e3                                          # test_krdem1.f90:160-166
    assert len(roots) == 3
    for t, jr in roots:
        if jr[0] != 0:
            errt = t - 2.5
        elif t <= 2.5:
            errt = t - 2.47
        else:
            errt = t - 2.53
        assert abs(errt) < 1e-3, (t, jr)                      # :180-192
    assert [bool(j[0]) for _, j in roots] == [False, True, False]


def check_krdem2(roots):
    assert len(roots) == 2
    for t, _ in roots:
        k = int(t / 81.2 + 0.5)
        assert abs(t - (81.17237787055 + (k - 1) * 81.41853556212)) < 1.0, t   # test_krdem2.f90:205-217
