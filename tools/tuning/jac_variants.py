"""A/B of the jac kernels: bitwise comparison of blocks/pivots + timing (scratch).  variant 8 = jac2, 0 = default"""
import ctypes as C, sys, numpy as np
sys.path.insert(0, '/root/repo')
import daskr_b200 as dk
from daskr_b200.api import DeviceBuffer, _ref
dk.init(0)
L = dk.lib()
np_, mx, my = 10, 1024, 1024
ns = 2 * np_
neq = ns * mx * my
iw = np.zeros(64, dtype=np.int32); dk.setup_rbdpre(mx, my, ns, np_, 0, iw); dk.web_setpar(np_, mx, my)
c, cdot = dk.web_cinit(neq, 1e5 * np_)
c *= 1 + 0.05 * np.random.default_rng(0).standard_normal(neq)
rewt = 1.0 / (1e-5 * np.abs(c) + 1e-5)
d_c, d_rewt, d_wk = DeviceBuffer.from_host(c), DeviceBuffer.from_host(rewt), DeviceBuffer(8 * neq)
nwp, niwp = L.dkb_rbd_lenwp(), L.dkb_rbd_leniwp()
ipar = np.array([1, 0], dtype=np.int32)
L.dkb_web_jac.argtypes = [C.c_void_p] * 16
res = {}
for var in (8, 0):
    dk.set_option(99, var)
    d_wp, d_iwp = DeviceBuffer(8 * nwp), DeviceBuffer(4 * niwp)
    neq_c, t, cjc, h, ier, ires = C.c_int(neq), C.c_double(0), C.c_double(3.7e3), C.c_double(0), C.c_int(0), C.c_int(0)
    L.dkb_web_jac(None, _ref(ires), _ref(neq_c), _ref(t), d_c.ptr, None, d_rewt.ptr, None, d_wk.ptr, _ref(h), _ref(cjc),
                  d_wp.ptr, d_iwp.ptr, _ref(ier), None, ipar.ctypes.data_as(C.c_void_p))
    L.dkb_sync()
    res[var] = (d_wp.to_host(np.float64, (nwp,)), d_iwp.to_host(np.int32, (niwp,)), ier.value)
    d_wp.free(); d_iwp.free()
print("ierr", res[8][2], res[0][2], "blocks equal", np.array_equal(res[8][0], res[0][0]), "pivots equal", np.array_equal(res[8][1], res[0][1]))
info = np.zeros(20, dtype=np.int32); info[[10, 11, 13, 14, 15]] = 1
iwork = np.zeros(40 + neq, dtype=np.int32); dk.setup_rbdpre(mx, my, ns, np_, 40, iwork); dk.web_setpar(np_, mx, my)
rwork = np.zeros(60); rpar = np.zeros(1)
c, cdot = dk.web_cinit(neq, 1e5 * np_)
dk.set_option(99, 0)
t, idid = dk.daskr(dk.WEB_RES, neq, 0.0, c, cdot, 1.0, info, 1e-5, 1e-5, rwork, iwork, rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)
ms = C.c_double(0)
for var in (8, 0, 8, 0):
    L.dkb_debug_time_jac(var, 5, C.byref(ms))
    print("jac variant", var, "ms %.3f" % ms.value, flush=True)
