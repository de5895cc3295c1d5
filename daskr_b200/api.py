"""ctypes binding of include/daskr_b200.h and the host-side mirror of the reference interface.

Nothing here computes: every call goes through the C ABI of
`_lib/libdaskr_b200.so` into CUDA kernels.  A missing library or a missing GPU
raises -- there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_lib", "libdaskr_b200.so")

OPT_DEVICE_IO, OPT_FUSED, OPT_MAX_STEPS, OPT_ASYNC_OUT = 1, 2, 3, 4


class DaskrError(RuntimeError):
    pass


_lib = None

RES_T = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)
JACK_T = C.CFUNCTYPE(None, *([C.c_void_p] * 16))
PSOL_T = C.CFUNCTYPE(None, *([C.c_void_p] * 15))

# every symbol include/daskr_b200.h declares (tests check the .so exports all of them)
EXPORTS = [
    "dkb_init", "dkb_finalize", "dkb_last_error", "dkb_version", "dkb_stream", "dkb_set_option", "dkb_sync", "dkb_output_wait",
    "dkb_comm_unique_id", "dkb_comm_init", "dkb_setup_rbdpre", "dkb_setup_rbgpre", "dkb_slab", "dkb_web_setpar", "dkb_web_setpar_ex",
    "dkb_web_res",
    "dkb_web_jac", "dkb_web_psol", "dkb_web_cinit", "dkb_heat_setpar", "dkb_heat_res", "dkb_heat_jac",
    "dkb_heat_psol", "dkb_heat_uinit", "dkb_web_rt", "dkb_heat_rt", "dkb_daskr",
    "dkb_state_bytes", "dkb_state_export", "dkb_state_import", "dkb_dslvk", "dkb_ddwnrm", "dkb_psol_rbdpre",
    "dkb_rbd_lenwp", "dkb_rbd_leniwp", "dkb_rbd_to_reference", "dkb_rbd_from_reference",
    "dkb_dgefa_batched", "dkb_dgesl_batched", "dkb_malloc", "dkb_free", "dkb_memcpy_h2d", "dkb_memcpy_d2h",
    "dkb_memcpy_d2d", "dkb_memset", "dkb_prof_enable", "dkb_prof_reset", "dkb_prof_get", "dkb_launch_count", "dkb_krylov_stats", "dkb_reorth_stats", "dkb_alloc_workspace", "dkb_workspace_get",
    "dkb_vcopy", "dkb_vscale", "dkb_vscale_to", "dkb_vaxpy", "dkb_lincomb", "dkb_ddot", "dkb_ewt_set_invert", "dkb_mask_vt",
    "dkb_jac_rbdpre", "dkb_psol_rbdpre_tiles", "dkb_interp", "dkb_predict", "dkb_newton_update", "dkb_phi_update", "dkb_error_norms",
]


def lib():
    """Load libdaskr_b200.so (built in-tree by `daskr_b200.build.build_cuda()`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise DaskrError(
            "CUDA library %s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(daskr_b200 has no CPU fallback)" % _LIB_PATH
        )
    L = C.CDLL(_LIB_PATH, mode=C.RTLD_GLOBAL)
    L.dkb_last_error.restype = C.c_char_p
    L.dkb_stream.restype = C.c_void_p
    L.dkb_ddwnrm.restype = C.c_double
    L.dkb_launch_count.restype = C.c_int64
    L.dkb_state_bytes.restype = C.c_int64
    L.dkb_state_export.argtypes = [C.c_void_p, C.c_int64]
    L.dkb_state_import.argtypes = [C.c_void_p, C.c_int64]
    L.dkb_set_option.argtypes = [C.c_int, C.c_int64]
    L.dkb_web_cinit.argtypes = [C.c_void_p, C.c_void_p, C.c_double]
    L.dkb_web_setpar_ex.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double]
    L.dkb_heat_uinit.argtypes = [C.c_void_p, C.c_void_p]
    L.dkb_dgefa_batched.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
    L.dkb_dgesl_batched.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p]
    L.dkb_malloc.argtypes = [C.c_void_p, C.c_int64]
    L.dkb_free.argtypes = [C.c_void_p]
    L.dkb_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    L.dkb_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    L.dkb_memcpy_d2d.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    L.dkb_memset.argtypes = [C.c_void_p, C.c_int, C.c_int64]
    L.dkb_comm_unique_id.argtypes = [C.c_void_p]
    L.dkb_comm_init.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.dkb_prof_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.dkb_rbd_lenwp.restype = C.c_int64
    L.dkb_rbd_leniwp.restype = C.c_int64
    L.dkb_rbd_to_reference.argtypes = [C.c_void_p] * 4
    L.dkb_rbd_from_reference.argtypes = [C.c_void_p] * 4
    L.dkb_alloc_workspace.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int64]
    V, D, I64 = C.c_void_p, C.c_double, C.c_int64
    L.dkb_workspace_get.argtypes = [C.c_int, V, V]
    L.dkb_vcopy.argtypes = [I64, V, V]
    L.dkb_vscale.argtypes = [I64, D, V]
    L.dkb_vscale_to.argtypes = [I64, D, V, V]
    L.dkb_vaxpy.argtypes = [I64, D, V, V]
    L.dkb_lincomb.argtypes = [I64, D, V, D, V, V]
    L.dkb_ddot.argtypes = [I64, V, V, V]
    L.dkb_ewt_set_invert.argtypes = [I64, C.c_int, V, V, V, V, V, V, V]
    L.dkb_mask_vt.argtypes = [I64, V, V, V]
    L.dkb_interp.argtypes = [I64, D, D, C.c_int, V, I64, V, V, V]
    L.dkb_predict.argtypes = [I64, C.c_int, V, I64, V, V, V]
    L.dkb_newton_update.argtypes = [I64, V, D, V, V, V]
    L.dkb_phi_update.argtypes = [I64, C.c_int, C.c_int, V, I64, V]
    L.dkb_error_norms.argtypes = [I64, C.c_int, V, V, V, V, V]
    L.dkb_daskr.restype = None
    L.dkb_daskr.argtypes = [C.c_void_p] * 21
    _lib = L
    return L


def _check(rc, what):
    if rc != 0:
        raise DaskrError("%s failed (%d): %s" % (what, rc, lib().dkb_last_error().decode()))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _ref(x):
    return C.c_void_p(C.addressof(x))


def _fn(name):
    return C.cast(getattr(lib(), name), C.c_void_p)


class _LazyFn:
    """Address of a built-in device callback (dkb_web_res ...), resolved at first use."""

    def __init__(self, name):
        self.name = name

    @property
    def ptr(self):
        return _fn(self.name)


WEB_RES, WEB_JAC, WEB_PSOL = _LazyFn("dkb_web_res"), _LazyFn("dkb_web_jac"), _LazyFn("dkb_web_psol")
HEAT_RES, HEAT_JAC, HEAT_PSOL = _LazyFn("dkb_heat_res"), _LazyFn("dkb_heat_jac"), _LazyFn("dkb_heat_psol")
WEB_RT, HEAT_RT = _LazyFn("dkb_web_rt"), _LazyFn("dkb_heat_rt")


def init(device=0):
    _check(lib().dkb_init(int(device)), "dkb_init")


def finalize():
    if _lib is not None:
        _lib.dkb_finalize()


def state_export():
    """Device-side solver state (y, yprime, PHI, WP/IWP, ID, vector tolerances, SAVEd scalars) as one uint8 array.
    Together with rwork / iwork / t this is a checkpoint (include/daskr_b200.h)."""
    n = int(lib().dkb_state_bytes())
    if n <= 0:
        raise DaskrError("no solver state to export")
    buf = np.empty(n, dtype=np.uint8)
    _check(lib().dkb_state_export(_p(buf), n), "dkb_state_export")
    return buf


def state_import(buf):
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    _check(lib().dkb_state_import(_p(buf), buf.size), "dkb_state_import")


def output_wait():
    """OPT_ASYNC_OUT: block until the y / yprime of the last return have arrived in the caller's (pinned) arrays."""
    _check(lib().dkb_output_wait(), "dkb_output_wait")


def set_option(opt, value):
    _check(lib().dkb_set_option(int(opt), int(value)), "dkb_set_option")


def comm_init_from_torch(rank, world_size, group=None):
    """Create the library's NCCL communicator; the 128-byte id travels over torch.distributed."""
    import torch
    import torch.distributed as dist

    L = lib()
    idbuf = np.zeros(128, dtype=np.uint8)
    if rank == 0:
        _check(L.dkb_comm_unique_id(_p(idbuf)), "dkb_comm_unique_id")
    t = torch.from_numpy(idbuf)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0, group=group)
    idbuf = t.cpu().numpy().copy()
    _check(L.dkb_comm_init(_p(idbuf), int(rank), int(world_size)), "dkb_comm_init")


def setup_rbdpre(mx, my, ns, nsd, lid, iwork):
    """setup_rbdpre(mx, my, ns, nsd, lid, iwork) -- src/daskr_rbdpre.f90:102-156 (DMSET2)."""
    assert iwork.dtype == np.int32
    a = [C.c_int(int(v)) for v in (mx, my, ns, nsd, lid)]
    lib().dkb_setup_rbdpre(*[C.byref(v) for v in a], _p(iwork))


def setup_rbgpre(mx, my, ns, nsd, nxg, nyg, lid, iwork):
    """setup_rbgpre(mx, my, ns, nsd, nxg, nyg, lid, iwork) -- src/daskr_rbgpre.f90:107-158 (DGSET2): block grouping."""
    assert iwork.dtype == np.int32
    a = [C.c_int(int(v)) for v in (mx, my, ns, nsd, nxg, nyg, lid)]
    lib().dkb_setup_rbgpre(*[C.byref(v) for v in a], _p(iwork))


def slab():
    jy0, my = C.c_int(0), C.c_int(0)
    _check(lib().dkb_slab(C.byref(jy0), C.byref(my)), "dkb_slab")
    return jy0.value, my.value


def web_setpar(np_, mx, my, ay=1.0):
    _check(lib().dkb_web_setpar_ex(int(np_), int(mx), int(my), float(ay)), "dkb_web_setpar")


def web_cinit(neq, pred_ic=1e5):
    c = np.zeros(neq)
    cdot = np.zeros(neq)
    _check(lib().dkb_web_cinit(_p(c), _p(cdot), float(pred_ic)), "dkb_web_cinit")
    return c, cdot


def heat_setpar(m):
    _check(lib().dkb_heat_setpar(int(m)), "dkb_heat_setpar")


def heat_uinit(neq):
    u = np.zeros(neq)
    ud = np.zeros(neq)
    _check(lib().dkb_heat_uinit(_p(u), _p(ud)), "dkb_heat_uinit")
    return u, ud


def _cbptr(cb):
    if cb is None:
        return C.c_void_p(0)
    if isinstance(cb, _LazyFn):
        return cb.ptr
    return C.cast(cb, C.c_void_p)


def _ptr_of(a):
    """numpy array (host) or integer device address."""
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return _p(a)


def daskr(res, neq, t, y, ydot, tout, info, rtol, atol, rwork, iwork, rpar, ipar, jac, psol, rt=None, nrt=0,
          jroot=None):
    """DASKR(RES,NEQ,T,Y,YPRIME,TOUT,INFO,RTOL,ATOL,IDID,RWORK,LRW,IWORK,LIW,RPAR,IPAR,JAC,PSOL,RT,NRT,JROOT)
    -- src/daskr.f:1-3, Krylov option.  y/ydot: float64 numpy arrays (host) or device addresses
    (with OPT_DEVICE_IO); info/iwork/ipar int32, rwork/rpar float64, modified in place.
    rt: root routine (dkb_rt_t: device y/yprime in, nrt host values out) with nrt > 0; jroot: int32[nrt],
    filled when idid == 5.  rwork needs 60 + 3*nrt words.  Returns (t, idid)."""
    assert info.dtype == np.int32 and iwork.dtype == np.int32 and rwork.dtype == np.float64
    neq_c, nrt_c, idid_c = C.c_int(int(neq)), C.c_int(int(nrt)), C.c_int(0)
    t_c, tout_c = C.c_double(float(t)), C.c_double(float(tout))
    rtol_a = np.atleast_1d(np.asarray(rtol, dtype=np.float64))
    atol_a = np.atleast_1d(np.asarray(atol, dtype=np.float64))
    lrw, liw = C.c_int64(rwork.size), C.c_int64(iwork.size)
    if jroot is None:
        jroot = np.zeros(max(int(nrt), 1), dtype=np.int32)
    assert jroot.dtype == np.int32 and jroot.size >= nrt
    rp = _p(rpar) if rpar is not None else C.c_void_p(0)
    ip = _p(ipar) if ipar is not None else C.c_void_p(0)
    lib().dkb_daskr(
        _cbptr(res), _ref(neq_c), _ref(t_c), _ptr_of(y), _ptr_of(ydot),
        _ref(tout_c), _p(info), _p(rtol_a), _p(atol_a), _ref(idid_c),
        _p(rwork), _ref(lrw), _p(iwork), _ref(liw), rp, ip, _cbptr(jac),
        _cbptr(psol), _cbptr(rt) if rt is not None else C.c_void_p(0), _ref(nrt_c), _p(jroot),
    )
    if idid_c.value == -33:
        raise DaskrError("daskr: illegal input (idid=-33): %s" % lib().dkb_last_error().decode())
    if np.ndim(rtol) > 0:
        rtol[...] = rtol_a
        atol[...] = atol_a
    return t_c.value, idid_c.value


# ---------------------------------------------------------------- batched LINPACK micro-API
class DeviceBuffer:
    def __init__(self, nbytes):
        self.ptr = C.c_void_p(0)
        self.nbytes = int(nbytes)
        _check(lib().dkb_malloc(C.byref(self.ptr), self.nbytes), "dkb_malloc")

    @classmethod
    def from_host(cls, a):
        a = np.ascontiguousarray(a)
        b = cls(a.nbytes)
        _check(lib().dkb_memcpy_h2d(b.ptr, _p(a), a.nbytes), "dkb_memcpy_h2d")
        return b

    def to_host(self, dtype, shape):
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes <= self.nbytes
        _check(lib().dkb_memcpy_d2h(_p(out), self.ptr, out.nbytes), "dkb_memcpy_d2h")
        return out

    def free(self):
        if self.ptr:
            lib().dkb_free(self.ptr)
            self.ptr = C.c_void_p(0)

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def dgefa_batched(bd, ns):
    """Factor nblk blocks (host array bd[nblk, ns*ns], column-major blocks as in jac_rbdpre).
    Returns (lu, ipvt[nblk, ns] 1-based, info[nblk], first_failing_block)."""
    bd = np.ascontiguousarray(bd, dtype=np.float64)
    nblk = bd.size // (ns * ns)
    d_bd = DeviceBuffer.from_host(bd)
    d_ip = DeviceBuffer(4 * ns * nblk)
    d_info = DeviceBuffer(4 * nblk)
    first = C.c_int64(0)
    _check(lib().dkb_dgefa_batched(d_bd.ptr, d_ip.ptr, ns, nblk, d_info.ptr, C.byref(first)), "dkb_dgefa_batched")
    lu = d_bd.to_host(np.float64, (nblk, ns * ns))
    ip = d_ip.to_host(np.int32, (nblk, ns))
    info = d_info.to_host(np.int32, (nblk,))
    for b in (d_bd, d_ip, d_info):
        b.free()
    return lu, ip, info, first.value


def dgesl_batched(lu, ipvt, b, ns):
    lu = np.ascontiguousarray(lu, dtype=np.float64)
    ipvt = np.ascontiguousarray(ipvt, dtype=np.int32)
    b = np.ascontiguousarray(b, dtype=np.float64)
    nblk = lu.size // (ns * ns)
    d_lu, d_ip, d_b = DeviceBuffer.from_host(lu), DeviceBuffer.from_host(ipvt), DeviceBuffer.from_host(b)
    _check(lib().dkb_dgesl_batched(d_lu.ptr, d_ip.ptr, ns, nblk, d_b.ptr), "dkb_dgesl_batched")
    x = d_b.to_host(np.float64, b.shape)
    for buf in (d_lu, d_ip, d_b):
        buf.free()
    return x


# ---------------------------------------------------------------- profiling hooks
def prof_enable(on=True):
    _check(lib().dkb_prof_enable(1 if on else 0), "dkb_prof_enable")


def prof_reset():
    _check(lib().dkb_prof_reset(), "dkb_prof_reset")


def prof_get():
    names = C.create_string_buffer(48 * 32)
    ms = np.zeros(32)
    n = np.zeros(32, dtype=np.int64)
    k = lib().dkb_prof_get(names, _p(ms), _p(n), 32)
    out = {}
    for i in range(k):
        nm = names.raw[48 * i: 48 * i + 48].split(b"\0")[0].decode()
        out[nm] = {"ms": float(ms[i]), "count": int(n[i])}
    return out


def launch_count():
    return int(lib().dkb_launch_count())


def krylov_stats(n=16):
    """Krylov iterations since prof_reset() by basis index ll = 1..n."""
    h = np.zeros(n, dtype=np.int64)
    _check(lib().dkb_krylov_stats(_p(h), n), "dkb_krylov_stats")
    return h


# ---------------------------------------------------------------- whole-run drivers
def solve_web(np_, mx, my, touts, rtol=1e-5, atol=1e-5, jpre=1, max_steps=0, fused=True, ic=True, pred_ic=None,
              y_out=True, roots=False, nxg=0, nyg=0):
    """The food-web example as example/example_web.f90:884-954 runs it (Krylov, no grouping, nrt=0):
    one DASKR call for consistent initial values (info(11)=1, info(14)=1 -> idid=4), then integration
    with outputs at `touts`.  Returns dict(y=[nout, neq], counters=[nout, 40], rhead=[nout, 60], idid, t).
    With max_steps > 0 the run stops after that many BDF steps (info(3)=1 stepping).  roots=True adds the
    example's root function (nrt=1, average(c1) - 20, example_web.f90:600-617); every idid=5 return is
    recorded in result["roots"] as (t, jroot, counters) and the integration continues.  nxg, nyg > 0: block
    grouping (jbg = 1, nxg x nyg groups; the example's METH = 2 uses 5 x 5, example_web.f90:873-877)."""
    ns = 2 * np_
    iwork = np.zeros(40 + 1, dtype=np.int32)  # sized below once the slab is known
    probe = np.zeros(64, dtype=np.int32)
    setup_rbdpre(mx, my, ns, np_, 0, probe)
    _, myloc = slab()
    neq = ns * mx * myloc
    iwork = np.zeros(40 + neq, dtype=np.int32)
    jbg = 1 if nxg > 0 else 0
    if jbg:
        setup_rbgpre(mx, my, ns, np_, nxg, nyg, 40, iwork)
    else:
        setup_rbdpre(mx, my, ns, np_, 40, iwork)
    web_setpar(np_, mx, my)
    set_option(OPT_FUSED, 2 if fused is True else int(fused))   # 0 callbacks, 1 two kernels, 2 single fused kernel
    info = np.zeros(20, dtype=np.int32)
    info[11] = 1  # info(12): Krylov
    info[14] = 1  # info(15): jac supplied
    info[15] = 1  # info(16): error test on differential variables only
    nrt = 1 if roots else 0
    rtf = WEB_RT if roots else None
    jroot = np.zeros(1, dtype=np.int32)
    found = []
    rwork = np.zeros(60 + 3 * nrt)
    rpar = np.zeros(1)
    ipar = np.array([jpre, jbg], dtype=np.int32)
    if pred_ic is None:
        pred_ic = 1e5 * np_   # example_web.f90:784 uses 1e5 for np=1 (= ee * prey level * np): start near the non-trivial branch
    c, cdot = web_cinit(neq, pred_ic)
    t = 0.0
    idid = 0
    if ic:
        info[10] = 1  # info(11)
        info[13] = 1  # info(14)
        t, idid = daskr(WEB_RES, neq, t, c, cdot, touts[0], info, rtol, atol, rwork, iwork, rpar, ipar, WEB_JAC, WEB_PSOL,
                        rtf, nrt, jroot)
        if idid < 0:
            return dict(y=None, counters=iwork[:40].copy()[None], rhead=rwork[:60].copy()[None], idid=idid, t=t,
                        c0=c, cdot0=cdot)
        info[10] = 0
    c0, cdot0 = c.copy(), cdot.copy()
    if max_steps > 0:
        info[2] = 1  # info(3): return after every step
    ys, cnts, rhs = [], [], []
    steps = 0
    for tout in touts:
        while True:
            t, idid = daskr(WEB_RES, neq, t, c, cdot, tout, info, rtol, atol, rwork, iwork, rpar, ipar, WEB_JAC, WEB_PSOL,
                            rtf, nrt, jroot)
            if idid < 0:
                break
            if idid == 5:
                found.append((t, jroot.copy(), iwork[:40].copy()))
                continue
            if max_steps > 0 and idid == 1:
                steps += 1
                if steps >= max_steps:
                    break
                continue
            break
        if y_out:
            ys.append(c.copy())
        cnts.append(iwork[:40].copy())
        rhs.append(rwork[:60].copy())
        if idid < 0 or (max_steps > 0 and steps >= max_steps):
            break
    return dict(y=np.array(ys) if y_out else None, counters=np.array(cnts), rhead=np.array(rhs), idid=idid, t=t,
                c0=c0, cdot0=cdot0, roots=found)


def solve_heat(m, touts, rtol=0.0, atol=1e-5, max_steps=0, fused=True, y_out=True, roots=False):
    """example/example_heat.f90:174-240 with the rbdpre ns=1 diagonal preconditioner.  roots=True adds the
    example's two root functions (max(u) - 0.1, max(u) - 0.01, example_heat.f90:91-109)."""
    probe = np.zeros(64, dtype=np.int32)
    setup_rbdpre(m + 2, m + 2, 1, 1, 0, probe)
    _, myloc = slab()
    neq = (m + 2) * myloc
    iwork = np.zeros(40, dtype=np.int32)
    setup_rbdpre(m + 2, m + 2, 1, 1, 0, iwork)
    heat_setpar(m)
    set_option(OPT_FUSED, 2 if fused is True else int(fused))   # 0 callbacks, 1 two kernels, 2 single fused kernel
    info = np.zeros(20, dtype=np.int32)
    info[11] = 1
    info[14] = 1
    nrt = 2 if roots else 0
    rtf = HEAT_RT if roots else None
    jroot = np.zeros(2, dtype=np.int32)
    found = []
    rwork = np.zeros(60 + 3 * nrt)
    rpar = np.zeros(2)
    ipar = np.array([0, 0, neq, m], dtype=np.int32)
    u, ud = heat_uinit(neq)
    if max_steps > 0:
        info[2] = 1
    t = 0.0
    idid = 0
    ys, cnts, rhs = [], [], []
    steps = 0
    for tout in touts:
        while True:
            t, idid = daskr(HEAT_RES, neq, t, u, ud, tout, info, rtol, atol, rwork, iwork, rpar, ipar, HEAT_JAC, HEAT_PSOL,
                            rtf, nrt, jroot)
            if idid < 0:
                break
            if idid == 5:
                found.append((t, jroot.copy(), iwork[:40].copy()))
                continue
            if max_steps > 0 and idid == 1:
                steps += 1
                if steps >= max_steps:
                    break
                continue
            break
        if y_out:
            ys.append(u.copy())
        cnts.append(iwork[:40].copy())
        rhs.append(rwork[:60].copy())
        if idid < 0 or (max_steps > 0 and steps >= max_steps):
            break
    return dict(y=np.array(ys) if y_out else None, counters=np.array(cnts), rhead=np.array(rhs), idid=idid, t=t,
                roots=found)
