import ctypes as C, time, sys
import numpy as np
sys.path.insert(0, '.')
import daskr_b200 as dk
from daskr_b200.api import DeviceBuffer, _ref
dk.init(0)
L = dk.lib()
import os
M = int(os.environ.get('GS_MESH', '1024'))
np_, mx, my = 10, M, int(os.environ.get('GS_ROWS', M))
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
if os.environ.get('GS_VARIANT'): dk.set_option(99, int(os.environ['GS_VARIANT']))
for _ in range(3): call()
L.dkb_sync()
t0 = time.perf_counter()
for _ in range(20): call()
L.dkb_sync()
print("GS %dx%dx20 variant %s noedge %s: %%.3f ms per application" % (mx, my, os.environ.get("GS_VARIANT"), os.environ.get("DKB_GS_NOEDGE")) % ((time.perf_counter() - t0) / 20 * 1e3))
import ctypes
if hasattr(L, "dkb_debug_gs_clocks"):
    out = (ctypes.c_longlong * 2048)()
    call(); L.dkb_debug_gs_clocks(out)
    v = list(out)
    for t in range(0, 72, 6):
        w = v[t*8:t*8+6]
        if w[0]: print("t=%d: A %d  wait+halo+zin %d  B %d  sync %d  edge+C %d  total %d cycles" % (t, w[1]-w[0], w[2]-w[1], w[3]-w[2], w[4]-w[3], w[5]-w[4], w[5]-w[0]))

if hasattr(L, "dkb_debug_gss_clocks"):
    n = 4096 * 8
    out = (ctypes.c_longlong * n)()
    call(); L.dkb_sync(); L.dkb_debug_gss_clocks(out, n)
    v = np.array(list(out)).reshape(4096, 8)
    nb = (my + 52) // 8
    tot = v[10:nb-10]
    d = np.diff(tot[:, 0])
    f = lambda a, b: (tot[:, a] - tot[:, b]).mean()
    print("mean batch period %.0f clk; barrier %.0f | z loads issue %.0f | exchange loads issue %.0f | window->z stores %.0f | steps %.0f | z regs->smem %.0f | exchange poll %.0f"
          % (d.mean(), f(1, 0), f(6, 1), f(7, 6), f(2, 7), f(3, 2), f(4, 3), f(5, 4)))
