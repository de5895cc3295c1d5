"""Spatial Gauss-Seidel preconditioner factor on the device (SURVEY.md 8 f1;
gauss_seidel, example/example_web.f90:393-575) and the shipped product preconditioner jpre = 3.

The device sweep is a tile/wavefront reordering of the reference's lexicographic sweep that applies
the reference's expression to the reference's operands at every mesh point, so the comparison with
the oracle is BIT-EXACT, on meshes that are smaller than a tile, not multiples of the tile, and
several tiles in both directions.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gpu_psol(dk, np_, mx, my, jpre, cj, b, lu=None):
    """dkb_web_psol on device buffers, exactly as dspigm would call it."""
    from daskr_b200.api import DeviceBuffer, _ref

    L = dk.lib()
    ns = 2 * np_
    iw = np.zeros(64, dtype=np.int32)
    dk.setup_rbdpre(mx, my, ns, np_, 0, iw)
    dk.web_setpar(np_, mx, my)
    d_b, d_wk = DeviceBuffer.from_host(b), DeviceBuffer(b.nbytes)
    neq, t, cjc, eps, ierr = C.c_int(b.size), C.c_double(0), C.c_double(cj), C.c_double(0), C.c_int(0)
    ipar = np.array([jpre, 0], dtype=np.int32)
    L.dkb_web_psol(_ref(neq), _ref(t), None, None, None, d_wk.ptr, _ref(cjc), None, None, None, d_b.ptr, _ref(eps),
                   _ref(ierr), None, ipar.ctypes.data_as(C.c_void_p))
    assert L.dkb_sync() == 0 and ierr.value == 0
    return d_b.to_host(np.float64, b.shape)


@pytest.mark.parametrize("np_,mx,my", [(1, 20, 20), (1, 33, 31), (2, 70, 45), (3, 32, 64), (10, 64, 40), (10, 100, 70),
                                       (16, 40, 50),
                                       # strip edges of the streaming kernel: mx around multiples of 32 (the last strip holds
                                       # only the columns the later iterations shifted out), very few / many rows, one column
                                       (1, 28, 9), (1, 29, 9), (2, 31, 3), (2, 36, 2), (1, 60, 1), (5, 61, 140), (1, 1, 40),
                                       (1, 2, 2), (4, 96, 33), (7, 65, 17), (10, 300, 50), (1, 128, 700)])
@pytest.mark.parametrize("cj", [1e8, 2.5e3, 7.0])
@pytest.mark.parametrize("kernel", ["auto", "stream"])
def test_gauss_seidel_bit_exact(dk, oracle, np_, mx, my, cj, kernel, monkeypatch):
    """kernel = auto: the library picks the one-pass streaming kernel or the five-pass tile scheduler by slab shape;
    stream: the one-pass kernel on every shape (DKB_GS_STREAM=1).  The tile scheduler alone: test below."""
    if kernel == "stream":
        monkeypatch.setenv("DKB_GS_STREAM", "1")
    ns = 2 * np_
    rng = np.random.default_rng(mx * 1000 + my + np_)
    b = rng.standard_normal(ns * mx * my)
    got = _gpu_psol(dk, np_, mx, my, 2, cj, b)
    oracle.L.o_web_setpar(np_, mx, my)
    z, x = b.copy(), np.zeros_like(b)
    oracle.L.o_web_gauss_seidel(b.size, 1.0 / cj, z.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p))
    assert np.array_equal(got, z)


@pytest.mark.parametrize("env", [{"DKB_GS_TY": "32"}, {"DKB_GS_TY": "7"}, {"DKB_GS_PER_STEP": "1"},
                                 {"DKB_GS_PER_STEP": "1", "DKB_GS_TY": "32"}, {}])
def test_gauss_seidel_variants_bit_exact(dk, oracle, env, monkeypatch):
    """The five-pass tile scheduler (DKB_GS_STREAM=0: the round-1 kernel, kept as a cross-check of the one-pass streaming
    kernel) in its tile heights and launch schedules gives the same bits."""
    monkeypatch.setenv("DKB_GS_STREAM", "0")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for np_, mx, my in ((10, 100, 70), (2, 70, 45), (10, 64, 130)):
        ns = 2 * np_
        b = np.random.default_rng(3).standard_normal(ns * mx * my)
        got = _gpu_psol(dk, np_, mx, my, 2, 2.5e3, b)
        oracle.L.o_web_setpar(np_, mx, my)
        z, x = b.copy(), np.zeros_like(b)
        oracle.L.o_web_gauss_seidel(b.size, 1.0 / 2.5e3, z.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p))
        assert np.array_equal(got, z)


def test_gauss_seidel_is_an_approximate_inverse(dk):
    """Property at a size the oracle is not needed for: A_S * GS(b) ~ b, A_S = I - (1/cj) J_diffusion."""
    np_, mx, my, cj = 10, 256, 192, 4.0e5
    ns = 2 * np_
    rng = np.random.default_rng(5)
    b = rng.standard_normal(ns * mx * my)
    z = _gpu_psol(dk, np_, mx, my, 2, cj, b).reshape(my, mx, ns)
    dx, dy = 1.0 / (mx - 1), 1.0 / (my - 1)
    diff = np.array([1.0] * np_ + [0.05] * np_)
    cox, coy = diff / dx ** 2, diff / dy ** 2
    zp = np.pad(z, ((1, 1), (1, 1), (0, 0)), mode="reflect")
    lap = cox * (zp[1:-1, 2:] - 2 * z + zp[1:-1, :-2]) + coy * (zp[2:, 1:-1] - 2 * z + zp[:-2, 1:-1])
    r = (z - lap / cj) - b.reshape(my, mx, ns)
    assert np.linalg.norm(r) <= 0.2 * np.linalg.norm(b)     # 5 sweeps: a preconditioner, not a solver


WEB_TOUTS = [1e-8, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, 1.0]


@pytest.mark.parametrize("jpre,mx", [(3, 20), (3, 40), (4, 20), (3, 32), (3, 64)])
def test_solve_web_with_spatial_factor(dk, oracle, jpre, mx):
    """example_web.f90 as shipped uses jpre = 3 (A_S then A_R).  Same run on both sides."""
    g = dk.solve_web(1, mx, mx, WEB_TOUTS, jpre=jpre)
    o = oracle.run_web(1, mx, mx, WEB_TOUTS, jpre=jpre)
    assert g["idid"] == o["idid"] == 3
    w = 1.0 / (1e-5 * np.abs(o["y"]) + 1e-5)
    err = np.sqrt(np.mean(((g["y"] - o["y"]) * w) ** 2, axis=1))
    print("jpre=%d %dx%d wrms(diff)/tol max %.3e; NST/NLI gpu %s cpu %s" % (
        jpre, mx, mx, err.max(), g["counters"][-1][[10, 19]], o["counters"][-1][[10, 19]]))
    assert err.max() <= 10.0
    for k in range(len(WEB_TOUTS)):
        cg, co = g["counters"][k], o["counters"][k]
        if cg[14] == 0 and co[14] == 0 and cg[13] == co[13]:
            assert abs(int(cg[10]) - int(co[10])) <= 2
            assert abs(int(cg[19]) - int(co[19])) <= max(3, 0.02 * co[19])


def test_solve_web_ns20_spatial_factor(dk, oracle):
    touts = [1e-6, 1e-4, 1e-2]
    g = dk.solve_web(10, 48, 40, touts, jpre=3)
    o = oracle.run_web(10, 48, 40, touts, jpre=3)
    assert g["idid"] == o["idid"] == 3
    w = 1.0 / (1e-5 * np.abs(o["y"]) + 1e-5)
    assert np.sqrt(np.mean(((g["y"] - o["y"]) * w) ** 2, axis=1)).max() <= 10.0
    assert np.abs(g["counters"][-1][[10, 19]].astype(int) - o["counters"][-1][[10, 19]].astype(int)).max() <= 3


@pytest.mark.parametrize("np_,mx,my,jpre", [(3, 16, 12, 3), (7, 12, 10, 3), (9, 32, 8, 3), (11, 10, 9, 1), (13, 9, 8, 3), (15, 8, 8, 1)])
def test_solve_web_any_block_size(dk, oracle, np_, mx, my, jpre):
    """The reference takes any ns (src/daskr_rbdpre.f90:102-156); so do residual, block Jacobian + LU, block solve and
    the Gauss-Seidel factor here (every ns = 2*np up to 32 is instantiated)."""
    touts = [1e-7, 1e-5, 1e-3]
    g = dk.solve_web(np_, mx, my, touts, jpre=jpre)
    o = oracle.run_web(np_, mx, my, touts, jpre=jpre)
    assert g["idid"] == o["idid"] == 3
    w = 1.0 / (1e-5 * np.abs(o["y"]) + 1e-5)
    err = np.sqrt(np.mean(((g["y"] - o["y"]) * w) ** 2, axis=1))
    print("np=%d %dx%d jpre=%d: wrms(diff)/tol %s; NST/NLI gpu %s cpu %s" % (np_, mx, my, jpre, err, g["counters"][-1][[10, 19]], o["counters"][-1][[10, 19]]))
    assert err.max() <= 10.0
    assert np.abs(g["counters"][-1][[10, 19]].astype(int) - o["counters"][-1][[10, 19]].astype(int)).max() <= 3


@pytest.mark.parametrize("np_,mx,my,pcols", [(1, 96, 40, 32), (2, 200, 33, 64), (10, 1024, 130, 512), (10, 2048, 64, 512), (3, 300, 300, 96)])
def test_column_panels_reproduce_the_sweep(dk, oracle, np_, mx, my, pcols, monkeypatch):
    """Several ranks sweep overlapped slabs AND overlapped column panels (32 columns of the left neighbour, 8 of the right
    one).  Same decay argument as for the slabs; checked here on one GPU (DKB_GS_FORCE_PANELS) against the exact sweep:
    the owned columns agree to rounding level on the unit square's mesh ratio, far below what a preconditioner needs."""
    ns = 2 * np_
    rng = np.random.default_rng(11)
    b = rng.standard_normal(ns * mx * my)
    for cj in (2.5e3, 40.0):
        exact = _gpu_psol(dk, np_, mx, my, 2, cj, b)
        monkeypatch.setenv("DKB_GS_FORCE_PANELS", "1")
        monkeypatch.setenv("DKB_GS_PANEL_COLS", str(pcols))
        cut = _gpu_psol(dk, np_, mx, my, 2, cj, b)
        monkeypatch.delenv("DKB_GS_FORCE_PANELS")
        monkeypatch.delenv("DKB_GS_PANEL_COLS")
        rel = np.abs(cut - exact).max() / np.abs(exact).max()
        print("%dx%d ns=%d cj=%g panels of %d columns: max |cut - exact| / max|exact| = %.2e" % (mx, my, ns, cj, pcols, rel))
        # (beta/(1-gamma))^32: ~(1/3)^32 = 5e-16 when dx = dy, up to ~(1/2)^32 = 2e-10 (times a modest constant) when dx << dy
        assert rel < (1e-12 if mx == my else 1e-7)
        assert not np.array_equal(cut, exact) or mx <= pcols * 1.5     # the panels were really used
