import ctypes as C, sys, numpy as np
sys.path.insert(0, '/root/repo')
import daskr_b200 as dk
dk.init(0)
import os
M = int(os.environ.get('JAC_MESH', '1024'))
np_, mx, my = 10, M, M
ns = 2*np_
probe = np.zeros(64, dtype=np.int32); dk.setup_rbdpre(mx, my, ns, np_, 0, probe)
neq = ns*mx*my
iwork = np.zeros(40+neq, dtype=np.int32); dk.setup_rbdpre(mx, my, ns, np_, 40, iwork); dk.web_setpar(np_, mx, my)
info = np.zeros(20, dtype=np.int32); info[[10,11,13,14,15]] = 1
rwork = np.zeros(60); rpar = np.zeros(1); ipar = np.array([int(os.environ.get('JAC_JPRE', '1')),0], dtype=np.int32)
c, cdot = dk.web_cinit(neq)
t, idid = dk.daskr(dk.WEB_RES, neq, 0.0, c, cdot, 1.0, info, 1e-5, 1e-5, rwork, iwork, rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)
L = dk.lib(); ms = C.c_double(0)
for var in [int(v) for v in os.environ.get('JAC_VARIANTS', '0,10,11,9,0').split(',')]:   # 0: jac3.cu; 10..15: jac2.cu; 9: one mapping
    rc = L.dkb_debug_time_jac(var, 5, C.byref(ms))
    print("jac variant", var, "rc", rc, "ms %.3f" % ms.value, flush=True)
