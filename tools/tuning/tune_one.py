import ctypes as C, sys, numpy as np
sys.path.insert(0, '/root/repo')
import daskr_b200 as dk
level, var, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dk.init(0)
np_, mx, my = 10, 1024, 1024
ns = 2*np_
probe = np.zeros(64, dtype=np.int32); dk.setup_rbdpre(mx, my, ns, np_, 0, probe)
neq = ns*mx*my
iwork = np.zeros(40+neq, dtype=np.int32); dk.setup_rbdpre(mx, my, ns, np_, 40, iwork); dk.web_setpar(np_, mx, my)
info = np.zeros(20, dtype=np.int32); info[[10,11,13,14,15]] = 1
rwork = np.zeros(60); rpar = np.zeros(1); ipar = np.array([1,0], dtype=np.int32)
c, cdot = dk.web_cinit(neq)
t, idid = dk.daskr(dk.WEB_RES, neq, 0.0, c, cdot, 1.0, info, 1e-5, 1e-5, rwork, iwork, rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)
info[10] = 0; dk.set_option(dk.OPT_MAX_STEPS, 2)
t, idid = dk.daskr(dk.WEB_RES, neq, t, c, cdot, 1.0, info, 1e-5, 1e-5, rwork, iwork, rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)
ms = C.c_double(0)
rc = dk.lib().dkb_debug_time_datv(level, var, reps, C.byref(ms))
print("level", level, "variant", var, "ms %.4f" % ms.value)
