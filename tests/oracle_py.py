"""ctypes wrapper over oracle/_build/liboracle.so -- TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference's Krylov path (see
oracle/oracle.h).  Only tests/, __graft_entry__.smoke() and bench.py's CPU
baseline legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "_build", "liboracle.so")

_lib = None


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def load():
    global _lib
    if _lib is not None:
        return Oracle(_lib)
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".cpp", ".h"))]
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.run(["make"], cwd=ORACLE_DIR, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    L = C.CDLL(LIB)
    L.o_ddot.restype = C.c_double
    L.o_dnrm2.restype = C.c_double
    L.o_ddwnrm.restype = C.c_double
    L.o_web_c1_average.restype = C.c_double
    L.o_ddot.argtypes = [C.c_long, C.c_void_p, C.c_void_p]
    L.o_dnrm2.argtypes = [C.c_long, C.c_void_p]
    L.o_ddwnrm.argtypes = [C.c_long, C.c_void_p, C.c_void_p]
    L.o_idamax.argtypes = [C.c_int, C.c_void_p]
    L.o_dgefa.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.o_dgesl.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.o_dgefa_batched.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_long, C.c_void_p]
    L.o_dgesl_batched.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_long, C.c_void_p]
    L.o_web_setpar.argtypes = [C.c_int, C.c_int, C.c_int]
    L.o_web_cinit.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    L.o_web_f.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    L.o_setup_rbdpre.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.o_jac_rbdpre.argtypes = [C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                               C.c_void_p, C.c_void_p, C.c_void_p]
    L.o_psol_rbdpre.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.o_heat_setpar.argtypes = [C.c_int]
    L.o_heat_uinit.argtypes = [C.c_void_p, C.c_void_p]
    L.o_run_web.restype = C.c_int
    L.o_run_web.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p,
                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_long, C.c_void_p]
    L.o_run_web_bg.restype = C.c_int
    L.o_run_web_bg.argtypes = [C.c_int, C.c_int] + list(L.o_run_web.argtypes)
    L.o_run_heat.restype = C.c_int
    L.o_run_heat.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_void_p, C.c_long, C.c_void_p]
    L.o_dslvk.argtypes = [C.c_long, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                          C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    L.o_web_gauss_seidel.argtypes = [C.c_long, C.c_double, C.c_void_p, C.c_void_p]
    _lib = L
    return Oracle(L)


class Oracle:
    def __init__(self, L):
        self.L = L

    # ---- BLAS / LINPACK ----
    def ddot(self, x, y):
        return self.L.o_ddot(x.size, _p(x), _p(y))

    def dnrm2(self, x):
        return self.L.o_dnrm2(x.size, _p(x))

    def ddwnrm(self, v, w):
        return self.L.o_ddwnrm(v.size, _p(v), _p(w))

    def idamax(self, x):
        return self.L.o_idamax(x.size, _p(x))

    def dgefa_batched(self, bd, ns):
        bd = np.array(bd, dtype=np.float64, order="C", copy=True)
        nblk = bd.size // (ns * ns)
        ip = np.zeros((nblk, ns), dtype=np.int32)
        info = np.zeros(nblk, dtype=np.int32)
        self.L.o_dgefa_batched(_p(bd), _p(ip), ns, nblk, _p(info))
        return bd.reshape(nblk, ns * ns), ip, info

    def dgesl_batched(self, lu, ip, b, ns):
        lu = np.ascontiguousarray(lu, dtype=np.float64)
        ip = np.ascontiguousarray(ip, dtype=np.int32)
        x = np.array(b, dtype=np.float64, order="C", copy=True)
        nblk = lu.size // (ns * ns)
        self.L.o_dgesl_batched(_p(lu), _p(ip), ns, nblk, _p(x))
        return x

    # ---- food web pieces ----
    def web_setup(self, np_, mx, my):
        self.L.o_web_setpar(np_, mx, my)
        iw = np.zeros(64, dtype=np.int32)
        self.L.o_setup_rbdpre(mx, my, 2 * np_, np_, 0, _p(iw))

    def web_cinit(self, np_, mx, my, pred_ic=1e5):
        neq = 2 * np_ * mx * my
        c, cdot, rpar = np.zeros(neq), np.zeros(neq), np.zeros(neq)
        self.L.o_web_cinit(_p(c), _p(cdot), pred_ic, _p(rpar))
        return c, cdot

    def web_res(self, c, cdot, t=0.0):
        delta, rpar = np.zeros_like(c), np.zeros_like(c)
        tt, cj, ires = C.c_double(t), C.c_double(0.0), C.c_int(0)
        ipar = np.zeros(2, dtype=np.int32)
        self.L.o_web_res(C.byref(tt), _p(c), _p(cdot), C.byref(cj), _p(delta), C.byref(ires), _p(rpar), _p(ipar))
        return delta, rpar

    def web_jac(self, c, rewt, cj, ns):
        """jac_rbdpre with rblock = rates and r0 = rates at c (what res leaves in rpar)."""
        _, r0 = self.web_res(c, np.zeros_like(c))
        u = c.copy()
        npts = c.size // ns
        bd = np.zeros(ns * ns * npts)
        ip = np.zeros(ns * npts, dtype=np.int32)
        r1 = np.zeros(ns)
        ierr = C.c_int(0)
        rates = C.cast(self.L.o_web_rates, C.c_void_p)
        self.L.o_jac_rbdpre(0.0, _p(u), _p(r0), rates, _p(r1), _p(rewt), cj, _p(bd), _p(ip), C.byref(ierr))
        assert np.array_equal(u, c)
        return bd.reshape(npts, ns * ns), ip.reshape(npts, ns), ierr.value

    def psol_rbdpre(self, b, bd, ip):
        x = b.copy()
        self.L.o_psol_rbdpre(_p(x), _p(np.ascontiguousarray(bd)), _p(np.ascontiguousarray(ip)))
        return x

    # ---- whole runs ----
    def run_web(self, np_, mx, my, touts, jpre=1, rtol=1e-5, atol=1e-5, max_steps=0, nxg=0, nyg=0):
        """nxg, nyg > 0: block grouping (jbg = 1, src/daskr_rbgpre.f90)."""
        neq = 2 * np_ * mx * my
        nout = len(touts)
        touts = np.asarray(touts, dtype=np.float64)
        y = np.zeros((nout, neq))
        cnt = np.zeros((nout, 40), dtype=np.int32)
        rh = np.zeros((nout, 60))
        tf = C.c_double(0)
        if nxg > 0:
            idid = self.L.o_run_web_bg(nxg, nyg, np_, mx, my, jpre, rtol, atol, nout, _p(touts), _p(y), _p(cnt), _p(rh),
                                       max_steps, C.byref(tf))
        else:
            idid = self.L.o_run_web(np_, mx, my, jpre, rtol, atol, nout, _p(touts), _p(y), _p(cnt), _p(rh), max_steps,
                                    C.byref(tf))
        return dict(y=y, counters=cnt, rhead=rh, idid=idid, t=tf.value)

    def run_heat(self, m, touts, rtol=0.0, atol=1e-5, max_steps=0):
        neq = (m + 2) ** 2
        nout = len(touts)
        touts = np.asarray(touts, dtype=np.float64)
        y = np.zeros((nout, neq))
        cnt = np.zeros((nout, 40), dtype=np.int32)
        rh = np.zeros((nout, 60))
        tf = C.c_double(0)
        idid = self.L.o_run_heat(m, rtol, atol, nout, _p(touts), _p(y), _p(cnt), _p(rh), max_steps, C.byref(tf))
        return dict(y=y, counters=cnt, rhead=rh, idid=idid, t=tf.value)
