"""Times the Gauss-Seidel factor and a short jpre=3 run at 1024x1024x20 (scratch)."""
import ctypes as C, time, sys
import numpy as np
sys.path.insert(0, '.')
import daskr_b200 as dk
from daskr_b200.api import DeviceBuffer, _ref
dk.init(0)
L = dk.lib()
np_, mx, my = 10, 1024, 1024
ns = 2 * np_
iw = np.zeros(64, dtype=np.int32)
dk.setup_rbdpre(mx, my, ns, np_, 0, iw)
dk.web_setpar(np_, mx, my)
b = np.random.default_rng(0).standard_normal(ns * mx * my)
d_b, d_wk = DeviceBuffer.from_host(b), DeviceBuffer(b.nbytes)
neq, t, cjc, eps, ierr = C.c_int(b.size), C.c_double(0), C.c_double(1e5), C.c_double(0), C.c_int(0)
ipar = np.array([2, 0], dtype=np.int32)
def call():
    L.dkb_web_psol(_ref(neq), _ref(t), None, None, None, d_wk.ptr, _ref(cjc), None, None, None, d_b.ptr, _ref(eps),
                   _ref(ierr), None, ipar.ctypes.data_as(C.c_void_p))
for _ in range(3): call()
L.dkb_sync()
t0 = time.perf_counter()
for _ in range(20): call()
L.dkb_sync()
print("GS 1024x1024x20: %.3f ms per application" % ((time.perf_counter() - t0) / 20 * 1e3))
for jpre in (1, 3):
    t0 = time.perf_counter()
    r = dk.solve_web(10, 1024, 1024, [1.0], jpre=jpre, max_steps=60, y_out=False)
    c = r["counters"][-1]
    print("jpre=%d 60 steps: %.2f s, t=%.3e idid=%d NST %d NRE %d NJE %d NETF %d NCFN %d NCFL %d NNI %d NLI %d NPS %d" % (
        jpre, time.perf_counter() - t0, r["t"], r["idid"], c[10], c[11], c[12], c[13], c[14], c[15], c[18], c[19], c[20]))
