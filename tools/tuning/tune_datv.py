import ctypes as C, sys, numpy as np
sys.path.insert(0, '/root/repo')
import daskr_b200 as dk
dk.init(0)
import os
M = int(os.environ.get('DATV_MESH', '1024'))
np_, mx, my = 10, M, M
ns = 2*np_
probe = np.zeros(64, dtype=np.int32); dk.setup_rbdpre(mx, my, ns, np_, 0, probe)
neq = ns*mx*my
iwork = np.zeros(40+neq, dtype=np.int32); dk.setup_rbdpre(mx, my, ns, np_, 40, iwork); dk.web_setpar(np_, mx, my)
info = np.zeros(20, dtype=np.int32); info[[10,11,13,14,15]] = 1
rwork = np.zeros(60); rpar = np.zeros(1); ipar = np.array([1,0], dtype=np.int32)
c, cdot = dk.web_cinit(neq)
t, idid = dk.daskr(dk.WEB_RES, neq, 0.0, c, cdot, 1.0, info, 1e-5, 1e-5, rwork, iwork, rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)
print("IC idid", idid)
info[10] = 0; dk.set_option(dk.OPT_MAX_STEPS, 3)
t, idid = dk.daskr(dk.WEB_RES, neq, t, c, cdot, 1.0, info, 1e-5, 1e-5, rwork, iwork, rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)
L = dk.lib()
ms = C.c_double(0)
for level, var in [tuple(int(x) for x in lv.split(':')) for lv in os.environ.get('DATV_VARIANTS', '1:0,2:0,2:1,2:2,2:3,1:0').split(',')]:
    rc = L.dkb_debug_time_datv(level, var, 20, C.byref(ms))
    print("level", level, "variant", var, "rc", rc, "ms %.4f" % ms.value, "GB/s %.0f" % (neq*212/ms.value/1e6), flush=True)
