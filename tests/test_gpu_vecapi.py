"""The vector-op layer of the C ABI (csrc/vecapi.cu): every whole-vector statement of DASKR's host
logic as one call on device arrays, against the statement written in numpy with the reference's operation order
(bit for bit: same multiply / add sequence, no FMA on either side) and, where the oracle exports it, against the oracle."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 100_003


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture()
def env(dk):
    from daskr_b200.api import DeviceBuffer

    dk.finalize()
    dk.init(0)
    return dk.lib(), DeviceBuffer


def test_blas_like_statements(env, oracle):
    L, DB = env
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal(N), rng.standard_normal(N)
    dx, dy, dz = DB.from_host(x), DB.from_host(y), DB(8 * N)
    assert L.dkb_vcopy(N, dx.ptr, dz.ptr) == 0
    assert np.array_equal(dz.to_host(np.float64, (N,)), x)
    assert L.dkb_vscale_to(N, 0.37, dx.ptr, dz.ptr) == 0
    assert np.array_equal(dz.to_host(np.float64, (N,)), 0.37 * x)
    assert L.dkb_vscale(N, -2.5, dz.ptr) == 0
    assert np.array_equal(dz.to_host(np.float64, (N,)), -2.5 * (0.37 * x))
    assert L.dkb_lincomb(N, 1.5, dx.ptr, -0.25, dy.ptr, dz.ptr) == 0
    assert np.array_equal(dz.to_host(np.float64, (N,)), 1.5 * x + -0.25 * y)
    assert L.dkb_vaxpy(N, 3.0, dx.ptr, dy.ptr) == 0
    assert np.array_equal(dy.to_host(np.float64, (N,)), y + 3.0 * x)
    r = C.c_double(0)
    assert L.dkb_ddot(N, dx.ptr, dx.ptr, C.byref(r)) == 0
    assert abs(r.value - oracle.ddot(x, x)) <= 1e-12 * r.value        # summation order differs, values do not


def test_weights_and_mask(env):
    L, DB = env
    rng = np.random.default_rng(1)
    y = rng.standard_normal(N) * 10.0 ** rng.integers(-3, 4, N)
    idv = np.where(rng.random(N) < 0.5, 1, -1).astype(np.int32)
    dy, dwt, dvt, did = DB.from_host(y), DB(8 * N), DB(8 * N), DB.from_host(idv)
    rt, at, ier = np.array([1e-5]), np.array([1e-7]), C.c_int(-1)
    assert L.dkb_ewt_set_invert(N, 0, _p(rt), _p(at), dy.ptr, dwt.ptr, dvt.ptr, did.ptr, C.byref(ier)) == 0
    wt = 1.0 / (1e-5 * np.abs(y) + 1e-7)                               # ddawts :739-748 then dinvwt :776
    assert ier.value == 0
    assert np.array_equal(dwt.to_host(np.float64, (N,)), wt)
    assert np.array_equal(dvt.to_host(np.float64, (N,)), np.maximum(idv, 0) * wt)
    # vector tolerances on the device; a zero weight is reported by its (1-based) index
    rtv, atv = np.full(N, 1e-4), np.full(N, 1e-6)
    rtv[4711] = atv[4711] = 0.0
    drt, dat = DB.from_host(rtv), DB.from_host(atv)
    assert L.dkb_ewt_set_invert(N, 1, drt.ptr, dat.ptr, dy.ptr, dwt.ptr, None, None, C.byref(ier)) == 0
    assert ier.value == 4712
    assert L.dkb_mask_vt(N, did.ptr, dwt.ptr, dvt.ptr) == 0
    got, w2 = dvt.to_host(np.float64, (N,)), dwt.to_host(np.float64, (N,))
    ok = np.arange(N) != 4711
    assert np.array_equal(got[ok], np.maximum(idv, 0)[ok] * w2[ok])


def test_predict_interp_phi_newton(env):
    L, DB = env
    rng = np.random.default_rng(2)
    ld = N + 29
    phi = rng.standard_normal((6, ld))
    dphi = DB.from_host(phi)
    dyo, dypo = DB(8 * N), DB(8 * N)
    for kp1 in (1, 2, 4, 6):
        gama = np.concatenate([[0.0], rng.standard_normal(5)])
        assert L.dkb_predict(N, kp1, dphi.ptr, ld, _p(gama), dyo.ptr, dypo.ptr) == 0
        yy, yp = phi[0, :N].copy(), np.zeros(N)                        # dnedk:2808-2813
        for j in range(1, kp1):
            yy = yy + 1.0 * phi[j, :N]
            yp = yp + gama[j] * phi[j, :N]
        assert np.array_equal(dyo.to_host(np.float64, (N,)), yy)
        assert np.array_equal(dypo.to_host(np.float64, (N,)), yp)
    for kold in (1, 3, 5):
        psi = np.cumsum(0.01 + 0.01 * rng.random(6))
        t, tout = 1.0, 0.993
        assert L.dkb_interp(N, t, tout, kold, dphi.ptr, ld, _p(psi), dyo.ptr, dypo.ptr) == 0
        koldp1, temp1 = kold + 1, tout - t                              # ddatrp :806-824, statement for statement
        yout, ypout = phi[0, :N].copy(), np.zeros(N)
        c, d, gam = 1.0, 0.0, temp1 / psi[0]
        for j in range(2, koldp1 + 1):
            d = d * gam + c / psi[j - 2]
            c = c * gam
            gam = (temp1 + psi[j - 2]) / psi[j - 1]
            yout = yout + c * phi[j - 1, :N]
            ypout = ypout + d * phi[j - 1, :N]
        assert np.array_equal(dyo.to_host(np.float64, (N,)), yout)
        assert np.array_equal(dypo.to_host(np.float64, (N,)), ypout)
    # phi update, ddstp:458-466
    e = rng.standard_normal(N)
    de = DB.from_host(e)
    for kp1, wk in ((2, 1), (5, 1), (6, 0), (3, 0)):
        ref = phi.copy()
        if wk:
            ref[kp1, :N] = e
        ref[kp1 - 1, :N] = ref[kp1 - 1, :N] + e
        for j in range(kp1 - 2, -1, -1):
            ref[j, :N] = ref[j, :N] + ref[j + 1, :N]
        d2 = DB.from_host(phi)
        assert L.dkb_phi_update(N, kp1, wk, d2.ptr, ld, de.ptr) == 0
        got = d2.to_host(np.float64, (6, ld))
        assert np.array_equal(got[:, :N], ref[:, :N])
        assert np.array_equal(got[:, N:], phi[:, N:])                   # nothing past the vector length is touched
    # Newton update, dnsk:2988-2990
    y, ee, yp, delta = (rng.standard_normal(N) for _ in range(4))
    dy, dee, dyp, dd = DB.from_host(y), DB.from_host(ee), DB.from_host(yp), DB.from_host(delta)
    cj = 123.456
    assert L.dkb_newton_update(N, dd.ptr, cj, dy.ptr, dee.ptr, dyp.ptr) == 0
    assert np.array_equal(dy.to_host(np.float64, (N,)), y - delta)
    assert np.array_equal(dee.to_host(np.float64, (N,)), ee - delta)
    assert np.array_equal(dyp.to_host(np.float64, (N,)), yp - cj * delta)


def test_error_norms_and_workspace(env, oracle):
    L, DB = env
    rng = np.random.default_rng(3)
    e, p1, p0 = (rng.standard_normal(N) for _ in range(3))
    w = 1.0 / (1e-5 * np.abs(rng.standard_normal(N)) + 1e-5)
    de, dp1, dp0, dw = DB.from_host(e), DB.from_host(p1), DB.from_host(p0), DB.from_host(w)
    out = np.zeros(3)
    assert L.dkb_error_norms(N, 3, de.ptr, dw.ptr, dp1.ptr, dp0.ptr, _p(out)) == 0
    ref = [oracle.ddwnrm(e, w), oracle.ddwnrm(p1 + e, w), oracle.ddwnrm(p0 + (p1 + e), w)]   # ddstp:360-378
    assert np.allclose(out, ref, rtol=1e-13, atol=0)
    # the workspace of a host that runs DASKR's own driver
    assert L.dkb_alloc_workspace(N, 5, 5, 0, 0) == 0
    ptrs = []
    for which in range(9):
        ptr, ld = C.c_void_p(0), C.c_int64(-1)
        assert L.dkb_workspace_get(which, C.byref(ptr), C.byref(ld)) == 0
        assert ptr.value
        ptrs.append((ptr.value, ld.value))
    assert ptrs[7][1] >= N and ptrs[8][1] >= N                           # column distances of phi and v
    assert len({p for p, _ in ptrs}) == 9
    assert L.dkb_workspace_get(99, C.byref(ptr), C.byref(ld)) != 0
