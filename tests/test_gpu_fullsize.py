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
