"""GPU tests at BASELINE.json's full sizes, through size-independent properties (the oracle would take
minutes to hours here): linear-solve residuals, bitwise agreement between independent kernel paths,
sampled comparison with the oracle, analytic decay.  All through the C ABI."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_lu_sweep_large(dk, oracle):
    """BASELINE config 5 (micro-sweep): ns in {2,4,8,16,32} and 20 at 1e5-1e6 blocks.  Pivots and
    factors bit-exact against the oracle on a random sample of blocks; A x = b to rounding on all."""
    from daskr_b200.api import DeviceBuffer

    L = dk.lib()
    rng = np.random.default_rng(0)
    for ns, nblk in ((2, 1_000_000), (4, 1_000_000), (8, 500_000), (16, 200_000), (20, 200_000), (32, 100_000)):
        bd = rng.standard_normal((nblk, ns * ns))
        rhs = np.random.default_rng(1).standard_normal((nblk, ns))
        lu, ip, info, first = dk.dgefa_batched(bd, ns)
        assert first == 0 and not info.any()
        x = dk.dgesl_batched(lu, ip, rhs, ns)
        # property on every block: residual of the solve
        a = bd.reshape(nblk, ns, ns).transpose(0, 2, 1)   # column-major blocks -> matrices
        r = np.einsum("pij,pj->pi", a, x) - rhs
        scale = np.abs(a).max(axis=(1, 2)) * np.abs(x).max(axis=1) + np.abs(rhs).max(axis=1)
        assert (np.abs(r).max(axis=1) / scale).max() < 1e-9
        # bit-exact against the oracle on a sample
        pick = rng.choice(nblk, size=2000, replace=False)
        lu_o, ip_o, info_o = oracle.dgefa_batched(bd[pick], ns)
        assert np.array_equal(ip[pick], ip_o)
        assert np.array_equal(lu[pick], lu_o)
        assert np.array_equal(x[pick], oracle.dgesl_batched(lu_o, ip_o, rhs[pick], ns))
        # the interleaved (thread-per-point) solve agrees bit for bit with the reference-layout one
        iw = np.zeros(64, dtype=np.int32)
        dk.setup_rbdpre(nblk // 100, 100, ns, ns, 0, iw)   # any mesh with nblk points
        d_bd, d_ip = DeviceBuffer.from_host(lu), DeviceBuffer.from_host(ip)
        d_wp, d_iwp = DeviceBuffer(8 * L.dkb_rbd_lenwp()), DeviceBuffer(4 * L.dkb_rbd_leniwp())
        assert L.dkb_rbd_from_reference(d_bd.ptr, d_ip.ptr, d_wp.ptr, d_iwp.ptr) == 0
        if ns > 1:
            from daskr_b200.api import _ref

            dk.web_setpar(ns // 2, nblk // 100, 100) if ns % 2 == 0 else None
            d_b = DeviceBuffer.from_host(rhs)
            neq, t, cj, eps, ierr = C.c_int(nblk * ns), C.c_double(0), C.c_double(1), C.c_double(0), C.c_int(0)
            ipar = np.array([1, 0], dtype=np.int32)
            L.dkb_web_psol(_ref(neq), _ref(t), None, None, None, None, _ref(cj), None, d_wp.ptr, d_iwp.ptr, d_b.ptr,
                           _ref(eps), _ref(ierr), None, ipar.ctypes.data_as(C.c_void_p))
            assert L.dkb_sync() == 0
            assert np.array_equal(d_b.to_host(np.float64, rhs.shape), x)


def test_web_1024_ns20_paths_agree(dk):
    """BASELINE config 2 (food web ns=20, 1024x1024): the single-kernel datv, the two-kernel datv and
    the callback form take the same steps; the first two are bitwise identical."""
    runs = {}
    for lvl in (2, 1):
        r = dk.solve_web(10, 1024, 1024, [1.0], max_steps=6, fused=lvl)
        assert r["idid"] >= 0
        runs[lvl] = r
    assert np.array_equal(runs[2]["y"], runs[1]["y"])
    assert np.array_equal(runs[2]["counters"], runs[1]["counters"])
    y = runs[2]["y"][-1].reshape(1024, 1024, 20)
    assert np.isfinite(y).all()
    pmin, qmin = y[:, :, :10].min(), y[:, :, 10:].min()
    assert pmin > 1.0 and qmin > 1e4, (pmin, qmin)                 # prey positive, predators on the non-trivial branch
    cnt = runs[2]["counters"][-1]
    assert cnt[10] == 6 and cnt[14] == 0 and cnt[15] == 0           # 6 steps, no convergence failures
    # symmetry property of the problem: the initial profile and the rate factor are symmetric under
    # x <-> y exchange only through alpha*x*y and sin*sin, both symmetric => c(jx,jy) == c(jy,jx)
    assert np.allclose(y, y.transpose(1, 0, 2), rtol=1e-6, atol=0)   # up to rounding amplified by the stiff operator


def test_web_256_ns20_vs_oracle(dk, oracle):
    """Same problem on a 256x256 mesh, first steps, against the oracle (takes ~10 s on the CPU)."""
    g = dk.solve_web(10, 256, 256, [1.0], max_steps=8)
    o = oracle.run_web(10, 256, 256, [1.0], max_steps=8)
    w = 1.0 / (1e-5 * np.abs(o["y"][-1]) + 1e-5)
    err = np.sqrt(np.mean(((g["y"][-1] - o["y"][-1]) * w) ** 2))
    print("web 256x256 ns=20, 8 steps: wrms(diff)/tol = %.3e; NLI gpu=%d cpu=%d" % (err, g["counters"][-1][19], o["counters"][-1][19]))
    assert err <= 10.0
    assert np.abs(g["counters"][-1][[10, 12, 18, 19]].astype(int) - o["counters"][-1][[10, 12, 18, 19]].astype(int)).max() <= 2


def test_heat_8192(dk):
    """BASELINE config 3: heat equation on the 8192x8192 mesh (m = 8190) with the rbdpre ns=1 diagonal
    preconditioner, Krylov path.  Property: Dirichlet rows stay 0, the solution stays within the
    maximum principle, and the fused and callback paths agree."""
    m = 8190
    touts = [1e-5]
    a = dk.solve_heat(m, touts, max_steps=5, fused=True)
    b = dk.solve_heat(m, touts, max_steps=5, fused=False)
    assert a["idid"] >= 0 and b["idid"] >= 0
    u = a["y"][-1].reshape(m + 2, m + 2)
    assert np.isfinite(u).all()
    assert np.abs(u[0]).max() == 0.0 and np.abs(u[-1]).max() == 0.0 and np.abs(u[:, 0]).max() == 0.0
    assert u.max() <= 1.0 + 1e-6 and u.min() >= -1e-6
    assert np.allclose(a["y"][-1], b["y"][-1], rtol=1e-9, atol=1e-12)
    assert np.allclose(u, u.T, rtol=1e-9, atol=1e-12)               # x <-> y symmetry of the problem


# ---------------------------------------------------------------- single calls at full size, bit for bit
def _gpu_web_calls(dk, c, cdot, rewt, cj, b, zgs, np_, mx, my):
    """res, jac (+ interleaved blocks / pivots), psol_rbdpre and the Gauss-Seidel factor, each ONE call through the C ABI."""
    from daskr_b200.api import DeviceBuffer, _ref

    L = dk.lib()
    ns = 2 * np_
    iw = np.zeros(64, dtype=np.int32)
    dk.setup_rbdpre(mx, my, ns, np_, 0, iw)
    dk.web_setpar(np_, mx, my)
    out = {}
    dc, dcd, dd = DeviceBuffer.from_host(c), DeviceBuffer.from_host(cdot), DeviceBuffer(c.nbytes)
    t, cj0, ires, ierr = C.c_double(0.0), C.c_double(0.0), C.c_int(0), C.c_int(0)
    L.dkb_web_res(_ref(t), dc.ptr, dcd.ptr, _ref(cj0), dd.ptr, _ref(ires), None, None)
    out["res"] = dd.to_host(np.float64, c.shape)
    dcd.free()
    dd.free()
    dw = DeviceBuffer.from_host(rewt)
    dwp, diwp = DeviceBuffer(8 * L.dkb_rbd_lenwp()), DeviceBuffer(4 * L.dkb_rbd_leniwp())
    neq, h, cjc, eps = C.c_int(c.size), C.c_double(0.0), C.c_double(cj), C.c_double(0.0)
    ipar = np.array([1, 0], dtype=np.int32)
    L.dkb_web_jac(None, _ref(ires), _ref(neq), _ref(t), dc.ptr, None, dw.ptr, None, None, _ref(h), _ref(cjc), dwp.ptr, diwp.ptr,
                  _ref(ierr), None, ipar.ctypes.data_as(C.c_void_p))
    out["ierr"] = ierr.value
    dw.free()
    db = DeviceBuffer.from_host(b)
    L.dkb_web_psol(_ref(neq), _ref(t), None, None, None, None, _ref(cjc), None, dwp.ptr, diwp.ptr, db.ptr, _ref(eps), _ref(ierr),
                   None, ipar.ctypes.data_as(C.c_void_p))
    assert L.dkb_sync() == 0 and ierr.value == 0
    out["psol"] = db.to_host(np.float64, b.shape)
    db.free()
    if zgs is not None:
        dz, dwk = DeviceBuffer.from_host(zgs), DeviceBuffer(zgs.nbytes)
        ipar2 = np.array([2, 0], dtype=np.int32)   # jpre = 2: the Gauss-Seidel factor alone
        L.dkb_web_psol(_ref(neq), _ref(t), None, None, None, dwk.ptr, _ref(cjc), None, dwp.ptr, diwp.ptr, dz.ptr, _ref(eps),
                       _ref(ierr), None, ipar2.ctypes.data_as(C.c_void_p))
        assert L.dkb_sync() == 0 and ierr.value == 0
        out["gs"] = dz.to_host(np.float64, zgs.shape)
        dz.free()
        dwk.free()
    out["wp"], out["iwp"] = dwp, diwp
    return out


def _tiles_to_blocks(raw_wp, raw_iwp, ns):
    """interleaved tiles (32 points each; element (i,j) of point p at ((p/32)*ns*ns + j*ns + i)*32 + p%32) -> bd / ipbd"""
    nt = raw_wp.size // (32 * ns * ns)
    bd = raw_wp.reshape(nt, ns * ns, 32).transpose(0, 2, 1).reshape(nt * 32, ns * ns)
    ip = raw_iwp.reshape(nt, ns, 32).transpose(0, 2, 1).reshape(nt * 32, ns)
    return bd, ip


def test_web_1024_ns20_single_calls_bit_exact(dk, oracle):
    """BASELINE configs[1] size (1024 x 1024 x 20 = 21 M unknowns): ONE residual, ONE jac_rbdpre (blocks and pivots),
    ONE psol_rbdpre and ONE Gauss-Seidel factor application, array_equal to the oracle (~15 s of CPU)."""
    np_, mx, my, ns = 10, 1024, 1024, 20
    oracle.web_setup(np_, mx, my)
    c, cdot = oracle.web_cinit(np_, mx, my, 1e5 * np_)
    rng = np.random.default_rng(3)
    c = c * (1.0 + 0.01 * rng.standard_normal(c.size))
    cdot = cdot + rng.standard_normal(c.size)
    rewt = 1.0 / (1e-5 * np.abs(c) + 1e-5)
    cj = 1.0 / 3.7e-4
    b = rng.standard_normal(c.size)
    zgs = rng.standard_normal(c.size)
    g = _gpu_web_calls(dk, c, cdot, rewt, cj, b, zgs, np_, mx, my)
    d_o, _ = oracle.web_res(c, cdot)
    assert np.array_equal(g["res"], d_o)
    del d_o
    bd_o, ip_o, ierr_o = oracle.web_jac(c, rewt, cj, ns)
    assert g["ierr"] == ierr_o == 0
    raw_wp = g["wp"].to_host(np.float64, (g["wp"].nbytes // 8,))
    raw_iwp = g["iwp"].to_host(np.int32, (g["iwp"].nbytes // 4,))
    bd_g, ip_g = _tiles_to_blocks(raw_wp, raw_iwp, ns)
    assert np.array_equal(ip_g[:mx * my], ip_o), "pivot indices must be bit-exact"
    assert np.array_equal(bd_g[:mx * my], bd_o)
    del bd_g, raw_wp
    assert np.array_equal(g["psol"], oracle.psol_rbdpre(b, bd_o, ip_o))
    del bd_o
    z, x = zgs.copy(), np.zeros_like(zgs)
    oracle.L.o_web_gauss_seidel(z.size, 1.0 / cj, z.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p))
    assert np.array_equal(g["gs"], z), "max rel diff %.3e" % np.max(np.abs(g["gs"] - z) / np.maximum(np.abs(z), 1e-300))


def test_web_4096_ns20_sampled_rows_bit_exact(dk, oracle):
    """BASELINE configs[3] size (4096 x 4096 x 20 = 335 M unknowns; the block array alone is 6.7 G doubles, past 32-bit
    indexing): residual, jac_rbdpre blocks / pivots and psol_rbdpre of the full-size GPU calls, compared bit for bit with
    the oracle on windows of rows at the bottom, in the middle and at the top of the mesh (o_web_window: the restated
    routines on a few rows with the whole mesh's spacing and coordinates)."""
    np_, mx, my, ns = 10, 4096, 4096, 20
    row = ns * mx
    iw = np.zeros(64, dtype=np.int32)
    dk.setup_rbdpre(mx, my, ns, np_, 0, iw)
    dk.web_setpar(np_, mx, my)
    c, cdot = dk.web_cinit(ns * mx * my, 1e5 * np_)
    rng = np.random.default_rng(4)
    nr = 6
    windows = [0, 2045, my - nr]
    for w0 in windows:                                   # perturb the sampled rows (+ neighbours) only: cheap, and enough
        lo, hi = max(w0 - 1, 0) * row, min(w0 + nr + 1, my) * row
        c[lo:hi] *= 1.0 + 0.01 * rng.standard_normal(hi - lo)
        cdot[lo:hi] += rng.standard_normal(hi - lo)
    rewt = 1.0 / (1e-5 * np.abs(c) + 1e-5)
    cj = 1.0 / 3.7e-4
    b = np.ones(c.size)
    for w0 in windows:
        b[w0 * row:(w0 + nr) * row] = rng.standard_normal(nr * row)
    g = _gpu_web_calls(dk, c, cdot, rewt, cj, b, None, np_, mx, my)
    assert g["ierr"] == 0
    L = dk.lib()
    for w0 in windows:
        oracle.L.o_web_setpar(np_, mx, my)
        oracle.L.o_web_window(nr, w0)
        iwo = np.zeros(64, dtype=np.int32)
        oracle.L.o_setup_rbdpre(mx, nr, ns, np_, 0, iwo.ctypes.data_as(C.c_void_p))
        sl = slice(w0 * row, (w0 + nr) * row)
        cw, cdw = c[sl].copy(), cdot[sl].copy()
        d_o, _ = oracle.web_res(cw, cdw)
        # rows whose stencil neighbours are inside the window, or are the mirrored global boundary
        r_lo = 0 if w0 == 0 else 1
        r_hi = nr if w0 + nr == my else nr - 1
        got = g["res"][sl]
        assert np.array_equal(got[r_lo * row:r_hi * row], d_o[r_lo * row:r_hi * row]), "residual rows %d..%d" % (w0 + r_lo, w0 + r_hi)
        bd_o, ip_o, ierr_o = oracle.web_jac(cw, rewt[sl].copy(), cj, ns)
        assert ierr_o == 0
        # the window's tiles of the interleaved block / pivot arrays (a mesh row is mx/32 whole tiles)
        t0, nt = w0 * (mx // 32), nr * (mx // 32)
        raw_wp = np.empty(nt * 32 * ns * ns)
        raw_iwp = np.empty(nt * 32 * ns, dtype=np.int32)
        assert L.dkb_memcpy_d2h(raw_wp.ctypes.data_as(C.c_void_p), C.c_void_p(g["wp"].ptr.value + 8 * t0 * 32 * ns * ns), raw_wp.nbytes) == 0
        assert L.dkb_memcpy_d2h(raw_iwp.ctypes.data_as(C.c_void_p), C.c_void_p(g["iwp"].ptr.value + 4 * t0 * 32 * ns), raw_iwp.nbytes) == 0
        bd_g, ip_g = _tiles_to_blocks(raw_wp, raw_iwp, ns)
        assert np.array_equal(ip_g, ip_o), "pivots, rows from %d" % w0
        assert np.array_equal(bd_g, bd_o), "LU blocks, rows from %d" % w0
        assert np.array_equal(g["psol"][sl], oracle.psol_rbdpre(b[sl].copy(), bd_o, ip_o)), "block solves, rows from %d" % w0
    oracle.L.o_web_setpar(np_, mx, my)   # leave no window behind


def test_web_4096_ic_needs_the_spatial_factor(dk):
    """Why bench.py's 4096 x 4096 solve runs the preconditioner as shipped by the example (jpre = 3): with the reaction
    block factor alone (jpre = 1, "drbdpre") DASKR's consistent-initialisation Newton gives up on this mesh
    (idid = -12) -- the reference algorithm's behaviour, not a defect of the port -- while it converges at 1024^2 and
    2048^2 (both preconditioners) and at 4096^2 with the Gauss-Seidel factor in front.  (The oracle cannot run this
    mesh: 115 GB of host memory.  At every mesh both sides can run, idid and counters agree: test_gpu_parity.py.)"""
    res = {}
    for m, jpre in ((2048, 1), (4096, 1), (4096, 3)):
        r = dk.solve_web(10, m, m, [1.0], jpre=jpre, max_steps=1, y_out=False)
        res[(m, jpre)] = r["idid"]
        print("mesh %d jpre %d: idid %d, counters nre/nje/nni/nli/ncfn/ncfl %s" % (m, jpre, r["idid"], r["counters"][-1][[11, 12, 18, 19, 14, 15]].tolist()))
    assert res[(2048, 1)] >= 0
    assert res[(4096, 3)] >= 0
    assert res[(4096, 1)] == -12


def test_web_4096_gauss_seidel_bit_exact(dk, oracle):
    """The Gauss-Seidel factor at BASELINE configs[3] size (4096 x 4096 x 20): one application of the streaming kernel
    -- 129 strips per species group, 645 tickets on 296 resident CTAs, the chain of strips two ticket rounds long --
    array_equal to the oracle's five lexicographic sweeps (example_web.f90:393-575; ~1 min of CPU)."""
    from daskr_b200.api import DeviceBuffer, _ref

    np_, mx, my, ns = 10, 4096, 4096, 20
    oracle.web_setup(np_, mx, my)
    iw = np.zeros(64, dtype=np.int32)
    dk.setup_rbdpre(mx, my, ns, np_, 0, iw)
    dk.web_setpar(np_, mx, my)
    n = ns * mx * my
    rng = np.random.default_rng(12)
    zgs = rng.standard_normal(n)
    cj = 1.0 / 4.2e-5                                    # the step size of the benchmark window
    L = dk.lib()
    dz, dwk, dummy = DeviceBuffer.from_host(zgs), DeviceBuffer(zgs.nbytes), DeviceBuffer(64)
    neq, t, cjc, eps, ierr = C.c_int(n), C.c_double(0.0), C.c_double(cj), C.c_double(0.0), C.c_int(0)
    ipar2 = np.array([2, 0], dtype=np.int32)             # jpre = 2: the Gauss-Seidel factor alone
    L.dkb_web_psol(_ref(neq), _ref(t), None, None, None, dwk.ptr, _ref(cjc), None, dummy.ptr, dummy.ptr, dz.ptr, _ref(eps),
                   _ref(ierr), None, ipar2.ctypes.data_as(C.c_void_p))
    assert L.dkb_sync() == 0 and ierr.value == 0
    got = dz.to_host(np.float64, zgs.shape)
    dz.free()
    dwk.free()
    x = np.zeros_like(zgs)
    oracle.L.o_web_gauss_seidel(n, 1.0 / cj, zgs.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p))
    del x
    assert np.array_equal(got, zgs), "max rel diff %.3e" % np.max(np.abs(got - zgs) / np.maximum(np.abs(zgs), 1e-300))
