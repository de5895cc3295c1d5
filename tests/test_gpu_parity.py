"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bars (BASELINE.json north star): LU pivots bit-exact, LU factors <= 1e-12 relative (they come
out bit-identical because the kernels use the reference's operation order without FMA),
solutions at every output time within 10x the tolerance in DASKR's weighted RMS norm, with the
Newton / Krylov counters reported side by side.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NS_LIST = list(range(1, 33))   # every block size up to one warp per block (the reference takes any ns)


def wrms(diff, yref, rtol, atol):
    w = 1.0 / (rtol * np.abs(yref) + atol)
    return float(np.sqrt(np.mean((diff * w) ** 2)))


# ---------------------------------------------------------------- batched LINPACK (config E)
@pytest.mark.parametrize("ns", NS_LIST)
def test_dgefa_dgesl_bit_exact(dk, oracle, ns):
    rng = np.random.default_rng(0)
    nblk = 1237  # not a multiple of any group count: exercises the ragged tail
    bd = rng.standard_normal((nblk, ns * ns))
    # edge cases: singular block (zero column), pivot ties, all-zero block, already triangular
    bd[3].reshape(ns, ns)[0 if ns == 1 else 1, :] = 0.0  # column-major: zero column 1
    if ns > 1:
        m = bd[5].reshape(ns, ns)
        m[0, :] = [1.0 if i % 2 == 0 else -1.0 for i in range(ns)]  # |a(i,1)| all equal -> first max wins
    bd[7] = 0.0
    lu_g, ip_g, info_g, first = dk.dgefa_batched(bd, ns)
    lu_o, ip_o, info_o = oracle.dgefa_batched(bd, ns)
    assert np.array_equal(ip_g, ip_o), "pivot indices must be bit-exact"
    assert np.array_equal(info_g, info_o)
    bad = np.nonzero(info_o)[0]
    assert first == (bad[0] + 1 if bad.size else 0)
    assert np.array_equal(lu_g, lu_o), "LU factors differ (max rel %.3e)" % np.max(
        np.abs(lu_g - lu_o) / np.maximum(np.abs(lu_o), 1e-300))
    # solve on the non-singular blocks
    ok = info_o == 0
    rhs = np.random.default_rng(1).standard_normal((nblk, ns))
    x_g = dk.dgesl_batched(lu_o[ok], ip_o[ok], rhs[ok], ns)
    x_o = oracle.dgesl_batched(lu_o[ok], ip_o[ok], rhs[ok], ns)
    assert np.array_equal(x_g, x_o)


def test_dgefa_matches_lapack_pivots(dk):
    """Independent pin: LINPACK first-max partial pivoting picks the same rows as LAPACK dgetrf."""
    from scipy.linalg import lu_factor

    rng = np.random.default_rng(3)
    ns, nblk = 8, 200
    bd = rng.standard_normal((nblk, ns * ns))
    _, ip_g, info, _ = dk.dgefa_batched(bd, ns)
    for p in range(nblk):
        a = bd[p].reshape(ns, ns).T  # column-major block -> matrix
        _, piv = lu_factor(a)
        assert np.array_equal(ip_g[p] - 1, piv)


# ---------------------------------------------------------------- food-web kernels
def _web_state(oracle, np_, mx, my, seed=0):
    oracle.web_setup(np_, mx, my)
    c, cdot = oracle.web_cinit(np_, mx, my)
    rng = np.random.default_rng(seed)
    c = c * (1.0 + 0.01 * rng.standard_normal(c.size))
    cdot = cdot + rng.standard_normal(c.size)
    return c, cdot


def _web_setup_gpu(dk, np_, mx, my):
    iw = np.zeros(64, dtype=np.int32)
    dk.setup_rbdpre(mx, my, 2 * np_, np_, 0, iw)
    dk.web_setpar(np_, mx, my)


@pytest.mark.parametrize("np_,mx,my", [(1, 20, 20), (1, 7, 5), (2, 9, 6), (4, 12, 10), (10, 16, 12), (16, 6, 5),
                                       (1, 32, 9), (10, 64, 13), (5, 32, 32), (8, 96, 7),    # mx % 32 == 0: staged single-kernel residual
                                       (3, 9, 7), (7, 32, 5), (9, 11, 6), (11, 64, 4), (13, 6, 6), (15, 33, 3)])   # every even ns
def test_web_res_bit_exact(dk, oracle, np_, mx, my):
    from daskr_b200.api import DeviceBuffer, _ref

    c, cdot = _web_state(oracle, np_, mx, my)
    d_o, _ = oracle.web_res(c, cdot)
    _web_setup_gpu(dk, np_, mx, my)
    dc, dcd, dd = DeviceBuffer.from_host(c), DeviceBuffer.from_host(cdot), DeviceBuffer(c.nbytes)
    t, cj, ires = C.c_double(0.0), C.c_double(0.0), C.c_int(0)
    dk.lib().dkb_web_res(_ref(t), dc.ptr, dcd.ptr, _ref(cj), dd.ptr, _ref(ires), None, None)
    d_g = dd.to_host(np.float64, c.shape)
    assert np.array_equal(d_g, d_o), "max abs diff %.3e" % np.max(np.abs(d_g - d_o))


@pytest.mark.parametrize("np_,mx,my", [(1, 20, 20), (2, 9, 6), (10, 16, 12), (16, 6, 5),
                                       (3, 8, 5), (5, 33, 4), (7, 9, 6), (9, 32, 3), (11, 6, 5), (13, 7, 4), (14, 5, 5), (15, 6, 3)])
def test_web_jac_psol_bit_exact(dk, oracle, np_, mx, my):
    from daskr_b200.api import DeviceBuffer, _ref

    ns = 2 * np_
    c, _ = _web_state(oracle, np_, mx, my, seed=1)
    rewt = 1.0 / (1e-5 * np.abs(c) + 1e-5)
    cj = 1.0 / 3.7e-4
    bd_o, ip_o, ierr_o = oracle.web_jac(c, rewt, cj, ns)
    _web_setup_gpu(dk, np_, mx, my)
    npts = mx * my
    L = dk.lib()
    dc, dw = DeviceBuffer.from_host(c), DeviceBuffer.from_host(rewt)
    # the jac callback writes rwp / iwp in the tile-interleaved order; convert for comparison
    dwp, diwp = DeviceBuffer(8 * L.dkb_rbd_lenwp()), DeviceBuffer(4 * L.dkb_rbd_leniwp())
    dbd, dip = DeviceBuffer(8 * ns * ns * npts), DeviceBuffer(4 * ns * npts)
    t, cjc, ires, ierr, neq, h = C.c_double(0.0), C.c_double(cj), C.c_int(0), C.c_int(0), C.c_int(c.size), C.c_double(0)
    ipar = np.array([1, 0], dtype=np.int32)
    L.dkb_web_jac(None, _ref(ires), _ref(neq), _ref(t), dc.ptr, None, dw.ptr, None, None, _ref(h), _ref(cjc),
                  dwp.ptr, diwp.ptr, _ref(ierr), None, ipar.ctypes.data_as(C.c_void_p))
    assert L.dkb_rbd_to_reference(dwp.ptr, diwp.ptr, dbd.ptr, dip.ptr) == 0
    bd_g = dbd.to_host(np.float64, (npts, ns * ns))
    ip_g = dip.to_host(np.int32, (npts, ns))
    assert ierr.value == ierr_o == 0
    assert np.array_equal(ip_g, ip_o), "pivot indices must be bit-exact"
    assert np.array_equal(bd_g, bd_o), "max rel diff %.3e" % np.max(np.abs(bd_g - bd_o) / np.maximum(np.abs(bd_o), 1e-300))
    # psol on those blocks: reference-layout entry point and the callback on the interleaved arrays
    b = np.random.default_rng(2).standard_normal(c.size)
    x_o = oracle.psol_rbdpre(b, bd_o, ip_o)
    db = DeviceBuffer.from_host(b)
    L.dkb_psol_rbdpre(db.ptr, dbd.ptr, dip.ptr)
    assert L.dkb_sync() == 0
    assert np.array_equal(db.to_host(np.float64, b.shape), x_o)
    db2 = DeviceBuffer.from_host(b)
    eps = C.c_double(0.0)
    L.dkb_web_psol(_ref(neq), _ref(t), None, None, None, None, _ref(cjc), None, dwp.ptr, diwp.ptr, db2.ptr, _ref(eps),
                   _ref(ierr), None, ipar.ctypes.data_as(C.c_void_p))
    assert L.dkb_sync() == 0 and ierr.value == 0
    assert np.array_equal(db2.to_host(np.float64, b.shape), x_o)
    # round trip reference -> interleaved -> reference
    dwp2, diwp2 = DeviceBuffer(8 * L.dkb_rbd_lenwp()), DeviceBuffer(4 * L.dkb_rbd_leniwp())
    dbd2, dip2 = DeviceBuffer(8 * ns * ns * npts), DeviceBuffer(4 * ns * npts)
    assert L.dkb_rbd_from_reference(dbd.ptr, dip.ptr, dwp2.ptr, diwp2.ptr) == 0
    assert L.dkb_rbd_to_reference(dwp2.ptr, diwp2.ptr, dbd2.ptr, dip2.ptr) == 0
    assert np.array_equal(dbd2.to_host(np.float64, (npts, ns * ns)), bd_o)
    assert np.array_equal(dip2.to_host(np.int32, (npts, ns)), ip_o)


def test_ddwnrm(dk, oracle):
    from daskr_b200.api import DeviceBuffer, _ref

    rng = np.random.default_rng(5)
    for n in (1, 2, 31, 1000, 123457):
        v = rng.standard_normal(n) * 10.0 ** rng.integers(-8, 8, n)
        w = 1.0 / (1e-5 * np.abs(v) + 1e-5)
        dv, dw = DeviceBuffer.from_host(v), DeviceBuffer.from_host(w)
        nn = C.c_int(n)
        g = dk.lib().dkb_ddwnrm(_ref(nn), dv.ptr, dw.ptr)
        o = oracle.ddwnrm(v, w)
        assert abs(g - o) <= 1e-13 * o
    z = np.zeros(100)
    dz = DeviceBuffer.from_host(z)
    nn = C.c_int(100)
    assert dk.lib().dkb_ddwnrm(_ref(nn), dz.ptr, dz.ptr) == 0.0  # exactly 0 for the zero vector (:859)


# ---------------------------------------------------------------- whole solves
WEB_TOUTS = [1e-8, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, 1.0]
COUNTER_NAMES = {10: "NST", 11: "NRE", 12: "NJE", 13: "NETF", 14: "NCFN", 15: "NCFL", 18: "NNI", 19: "NLI", 20: "NPS"}


def _report(tag, touts, g, o, rtol, atol):
    worst = 0.0
    for k in range(len(g["y"])):
        e = wrms(g["y"][k] - o["y"][k], o["y"][k], rtol, atol)
        worst = max(worst, e)
        cg = {COUNTER_NAMES[i]: int(g["counters"][k][i]) for i in COUNTER_NAMES}
        co = {COUNTER_NAMES[i]: int(o["counters"][k][i]) for i in COUNTER_NAMES}
        print("%s t=%.1e wrms(diff)=%.3e gpu=%s cpu=%s" % (tag, touts[k], e, cg, co))
    return worst


@pytest.mark.parametrize("mx,fused", [(20, True), (20, False), (40, True), (32, 2), (32, 1), (64, 2)])
def test_solve_web_parity(dk, oracle, mx, fused):
    """BASELINE config A: food web ns=2 as shipped (20x20 of the original, 40x40 of the modern file),
    reaction-factor preconditioner (jpre=1), IC calculation on, outputs 1e-8 .. 1."""
    rtol = atol = 1e-5
    g = dk.solve_web(1, mx, mx, WEB_TOUTS, rtol, atol, jpre=1, fused=fused)
    o = oracle.run_web(1, mx, mx, WEB_TOUTS, jpre=1, rtol=rtol, atol=atol)
    assert g["idid"] == o["idid"] == 3
    assert len(g["y"]) == len(WEB_TOUTS)
    worst = _report("web%dx%d fused=%s" % (mx, mx, fused), WEB_TOUTS, g, o, rtol, atol)
    assert worst <= 10.0, "solution differs from the CPU path by %.3g x tolerance" % worst
    # Counters: while neither side has had a convergence failure the two runs take the same
    # decisions and the counters agree to within an iteration or two (only the reduction order
    # differs).  Once Newton/linear failures start (t > 0.1 with the reaction-only
    # preconditioner) step-size retries amplify last-bit differences, so later counts are
    # reported, not asserted tightly.
    for k in range(len(WEB_TOUTS)):
        cg, co = g["counters"][k], o["counters"][k]
        if cg[14] == 0 and co[14] == 0 and cg[13] == co[13]:
            assert abs(int(cg[19]) - int(co[19])) <= max(3, 0.02 * co[19]), (WEB_TOUTS[k], cg[19], co[19])
            assert abs(int(cg[10]) - int(co[10])) <= 2
    nli_g, nli_o = g["counters"][-1][19], o["counters"][-1][19]
    assert abs(int(nli_g) - int(nli_o)) <= 0.5 * nli_o


def test_solve_web_ns20_short(dk, oracle):
    """ns=20 (np=10) on a small mesh, first outputs only: the size class of BASELINE config B."""
    rtol = atol = 1e-5
    touts = [1e-8, 1e-7, 1e-6, 1e-5, 1e-4]
    g = dk.solve_web(10, 12, 12, touts, rtol, atol)
    o = oracle.run_web(10, 12, 12, touts, rtol=rtol, atol=atol)
    assert g["idid"] == o["idid"] == 3
    assert _report("web12x12 ns=20", touts, g, o, rtol, atol) <= 10.0


@pytest.mark.parametrize("m,fused", [(10, True), (10, False), (30, True)])
def test_solve_heat_parity(dk, oracle, m, fused):
    touts = [0.01 * 2 ** n for n in range(8)]
    g = dk.solve_heat(m, touts, 0.0, 1e-5, fused=fused)
    o = oracle.run_heat(m, touts, 0.0, 1e-5)
    assert g["idid"] == o["idid"] == 3
    assert _report("heat m=%d fused=%s" % (m, fused), touts, g, o, 0.0, 1e-5) <= 10.0
    # analytic pin: the slowest mode decays like exp(-2 pi^2 t) (discrete rate within 2%)
    u = g["y"]
    rate = np.log(np.abs(u[2]).max() / np.abs(u[3]).max()) / (touts[3] - touts[2])
    assert abs(rate - 2 * np.pi ** 2) / (2 * np.pi ** 2) < 0.02


def test_fused_equals_unfused_bitwise_state(dk):
    """The fused datv kernel and the five-statement form around res()/psol() run the same
    arithmetic: the initial-condition call must leave identical y on both paths."""
    a = dk.solve_web(1, 20, 20, [1e-8], fused=True)
    b = dk.solve_web(1, 20, 20, [1e-8], fused=False)
    assert np.array_equal(a["c0"], b["c0"])
    assert np.array_equal(a["counters"][0], b["counters"][0])
    # 32-wide meshes take the single-kernel datv (level 2); it must equal the two-kernel path
    # (level 1) and the callback path (level 0) bit for bit, including after real time steps
    for np_, mx, my in ((1, 32, 32), (10, 32, 12), (4, 64, 9)):
        runs = [dk.solve_web(np_, mx, my, [1e-8, 1e-5, 1e-3], fused=lvl) for lvl in (2, 1, 0)]
        # levels 2 and 1 share every reduction kernel: elementwise-identical datv => identical runs
        assert np.array_equal(runs[0]["c0"], runs[1]["c0"])
        assert np.array_equal(runs[0]["y"], runs[1]["y"]), (np_, mx, my, np.abs(runs[0]["y"] - runs[1]["y"]).max())
        assert np.array_equal(runs[0]["counters"], runs[1]["counters"])
        # level 0 forms ||v1|| with a separate reduction kernel (different summation order), so it
        # agrees to rounding, not bitwise, once Krylov iterations have happened
        assert np.allclose(runs[0]["c0"], runs[2]["c0"], rtol=1e-9, atol=1e-12)
        assert np.allclose(runs[0]["y"], runs[2]["y"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("ns", [1, 2, 5, 20, 32])
def test_dgefa_dgesl_empty_and_single(dk, oracle, ns):
    """Empty batch (no launch, no error) and a batch of one block."""
    lu, ip, info, first = dk.dgefa_batched(np.zeros((0, ns * ns)), ns)
    assert lu.shape[0] == 0 and ip.shape[0] == 0 and first == 0
    x = dk.dgesl_batched(np.zeros((0, ns * ns)), np.zeros((0, ns), dtype=np.int32), np.zeros((0, ns)), ns)
    assert x.shape[0] == 0
    rng = np.random.default_rng(11)
    bd, rhs = rng.standard_normal((1, ns * ns)), rng.standard_normal((1, ns))
    lu_g, ip_g, info_g, first = dk.dgefa_batched(bd, ns)
    lu_o, ip_o, info_o = oracle.dgefa_batched(bd, ns)
    assert np.array_equal(lu_g, lu_o) and np.array_equal(ip_g, ip_o) and first == 0
    assert np.array_equal(dk.dgesl_batched(lu_g, ip_g, rhs, ns), oracle.dgesl_batched(lu_o, ip_o, rhs, ns))


def _web_jac_gpu(dk, c, rewt, cj, ns, npts, variant=0):
    """dkb_web_jac through the C ABI with the kernel selector `variant` (option 99: 0 = default, 10 = jac2.cu,
    9 = the one-mapping kernel, 16 = default kernel followed by the forced exact recomputation)."""
    from daskr_b200.api import DeviceBuffer, _ref

    L = dk.lib()
    dk.set_option(99, variant)
    try:
        dc, dw = DeviceBuffer.from_host(c), DeviceBuffer.from_host(rewt)
        dwp, diwp = DeviceBuffer(8 * L.dkb_rbd_lenwp()), DeviceBuffer(4 * L.dkb_rbd_leniwp())
        dbd, dip = DeviceBuffer(8 * ns * ns * npts), DeviceBuffer(4 * ns * npts)
        t, cjc, ires, ierr, neq, h = C.c_double(0.0), C.c_double(cj), C.c_int(0), C.c_int(0), C.c_int(c.size), C.c_double(0)
        ipar = np.array([1, 0], dtype=np.int32)
        L.dkb_web_jac(None, _ref(ires), _ref(neq), _ref(t), dc.ptr, None, dw.ptr, None, None, _ref(h), _ref(cjc),
                      dwp.ptr, diwp.ptr, _ref(ierr), None, ipar.ctypes.data_as(C.c_void_p))
        assert L.dkb_rbd_to_reference(dwp.ptr, diwp.ptr, dbd.ptr, dip.ptr) == 0
        return dbd.to_host(np.float64, (npts, ns * ns)), dip.to_host(np.int32, (npts, ns)), ierr.value
    finally:
        dk.set_option(99, 0)


@pytest.mark.parametrize("np_,mx,my", [(10, 16, 12), (10, 35, 7), (11, 6, 5), (12, 9, 4)])
def test_web_jac_kernel_variants_agree(dk, oracle, np_, mx, my):
    """ns = 20 / 22 / 24 run jac3.cu (two rows of a block per lane) by default: same blocks and pivots, bit for bit,
    as the oracle, as jac2.cu, as the one-mapping kernel and as the exact recomputation path."""
    ns = 2 * np_
    c, _ = _web_state(oracle, np_, mx, my, seed=3)
    rewt = 1.0 / (1e-5 * np.abs(c) + 1e-5)
    cj = 1.0 / 2.1e-3
    bd_o, ip_o, ierr_o = oracle.web_jac(c, rewt, cj, ns)
    assert ierr_o == 0
    _web_setup_gpu(dk, np_, mx, my)
    for variant in (0, 16, 10, 9):
        bd_g, ip_g, ierr_g = _web_jac_gpu(dk, c, rewt, cj, ns, mx * my, variant)
        assert ierr_g == 0, variant
        assert np.array_equal(ip_g, ip_o), variant
        assert np.array_equal(bd_g, bd_o), variant
        assert np.array_equal(np.signbit(bd_g), np.signbit(bd_o)), variant   # zeros keep their sign too


@pytest.mark.parametrize("pt", [0, 37])
def test_web_jac_zero_pivot_is_redone_exactly(dk, oracle, pt):
    """A block with an exactly zero first pivot: the state of one mesh point is zero (the block is then diagonal) and
    cj cancels its first diagonal entry to the last bit.  jac3.cu only detects that; the answer must be LINPACK's
    (linpack.f90:394-397: info = k, the column is skipped, the factorisation goes on) from the exact kernel."""
    np_, mx, my = 10, 8, 6
    ns = 2 * np_
    c, _ = _web_state(oracle, np_, mx, my, seed=5)
    rewt = 1.0 / (1e-5 * np.abs(c) + 1e-5)
    c[pt * ns:(pt + 1) * ns] = 0.0
    rewt[pt * ns:(pt + 1) * ns] = 10.24
    # the unshifted first diagonal entry exactly as jac_rbdpre forms it: from the oracle with cj = 0
    bd0, _, _ = oracle.web_jac(c, rewt, 0.0, ns)
    cj = -bd0[pt, 0]
    bd_o, ip_o, ierr_o = oracle.web_jac(c, rewt, cj, ns)
    assert ierr_o != 0
    _web_setup_gpu(dk, np_, mx, my)
    bd_g, ip_g, ierr_g = _web_jac_gpu(dk, c, rewt, cj, ns, mx * my, 0)
    bd_x, ip_x, ierr_x = _web_jac_gpu(dk, c, rewt, cj, ns, mx * my, 10)
    assert ierr_g != 0 and ierr_g == ierr_x
    assert np.array_equal(ip_g, ip_x) and np.array_equal(bd_g, bd_x)
    # the reference stops at the first singular block (daskr_rbdpre.f90:246); up to and including it the factors agree
    assert np.array_equal(bd_g[:pt + 1], bd_o[:pt + 1]) and np.array_equal(ip_g[:pt + 1], ip_o[:pt + 1])


def test_solve_web_against_an_independent_integrator(dk, oracle):
    """dkb_daskr on the food web (ns = 2, 8 x 8) against tests/independent.py -- implicit Euler on the DAE with a dense
    finite-difference Jacobian, extrapolated twice; it shares only the right-hand side with DASKR.  Started from the
    GPU run's own consistent initial values (output at t = 1e-9)."""
    from tests.independent import extrapolated_implicit_euler

    np_, mx, my = 1, 8, 8
    ns = 2 * np_
    touts = [1e-9, 1e-3, 1e-1, 1.0]
    for jpre in (1, 3):
        g = dk.solve_web(np_, mx, my, touts, 1e-8, 1e-8, jpre=jpre)
        assert g["idid"] == 3
        if jpre == 1:
            oracle.web_setup(np_, mx, my)
            c0 = g["y"][0].copy()
            pred = (np.arange(c0.size) % ns) >= np_

            def f(c):
                d, _ = oracle.web_res(c, np.zeros_like(c))
                return -d

            best = extrapolated_implicit_euler(f, np.where(pred, 0.0, 1.0), c0, touts[0], touts[1:])
        err = np.max(np.abs(np.asarray(g["y"])[1:] - best) / np.abs(best), axis=1)
        print("jpre=%d: GPU vs extrapolated implicit Euler at t = 1e-3, 0.1, 1:" % jpre, err)
        assert err[0] <= 1e-5 and err[1] <= 1e-5 and err[2] <= 1e-7
