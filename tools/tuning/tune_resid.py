"""Times the residual half of datv (jpre = 3) for the kernel variants of web_resid_single (datv_fused.cu)."""
import ctypes as C, os, sys, numpy as np
sys.path.insert(0, '.')
import daskr_b200 as dk
dk.init(0)
M = int(os.environ.get('RES_MESH', '2048'))
np_, mx, my = 10, M, M
ns = 2 * np_
probe = np.zeros(64, dtype=np.int32); dk.setup_rbdpre(mx, my, ns, np_, 0, probe)
neq = ns * mx * my
iwork = np.zeros(40 + neq, dtype=np.int32); dk.setup_rbdpre(mx, my, ns, np_, 40, iwork); dk.web_setpar(np_, mx, my)
info = np.zeros(20, dtype=np.int32); info[[10, 11, 13, 14, 15]] = 1
rwork = np.zeros(60); rpar = np.zeros(1); ipar = np.array([3, 0], dtype=np.int32)
c, cdot = dk.web_cinit(neq, 1e5 * np_)
t, idid = dk.daskr(dk.WEB_RES, neq, 0.0, c, cdot, 1.0, info, 1e-5, 1e-5, rwork, iwork, rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)
print("IC idid", idid, flush=True)
L = dk.lib(); ms = C.c_double(0)
for var in [int(v) for v in os.environ.get('RES_VARIANTS', '0,21,22,23,0').split(',')]:
    rc = L.dkb_debug_time_resid(var, 20, C.byref(ms))
    print("resid variant %d rc %d  %.3f ms  %.0f GB/s (48 B per unknown)" % (var, rc, ms.value, neq * 48 / ms.value / 1e6), flush=True)
