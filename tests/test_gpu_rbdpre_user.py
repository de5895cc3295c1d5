"""jac_rbdpre with a USER reaction term (src/daskr_rbdpre.f90:158-251; hook rblock :176-189) through dkb_jac_rbdpre.

The user routine here is the food-web rate law (example_web.f90:271-299) written with torch elementwise operations in the
reference's operation order (separate multiply and add, accumulation over i = 1..ns), called back from the library with
device pointers -- i.e. NOT the library's built-in fused kernel.  Blocks and pivots must equal the oracle's jac_rbdpre
with rblock = rates bit for bit, psol_rbdpre on them as well, and the failure code must be dgefa's info of the first
singular block.  A second, made-up reaction term checks that nothing food-web-specific is left in the path.
"""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RBLOCK_T = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)


class _DevArray:
    """numpy-style view of device memory for torch.as_tensor (CUDA array interface)"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 3}


def _torch_view(ptr, n):
    import torch

    return torch.as_tensor(_DevArray(ptr, n), device="cuda")


def _web_rates_torch(np_, mx, my):
    """R(c) for all mesh points, example_web.f90:286-297, elementwise on the GPU in the reference's order."""
    import torch

    ns = 2 * np_
    aa, ee, gg, bb, alpha, beta = 1.0, 1e4, 0.5e-6, 1.0, 50.0, 300.0
    acoef = np.zeros((ns, ns))
    bcoef = np.zeros(ns)
    for j in range(np_):
        for i in range(np_):
            acoef[np_ + i, j] = ee
            acoef[i, np_ + j] = -gg
        acoef[j, j] = -aa
        acoef[np_ + j, np_ + j] = -aa
        bcoef[j], bcoef[np_ + j] = bb, -bb
    dx, dy = 1.0 / (mx - 1), 1.0 / (my - 1)
    pi = 4 * np.arctan(1.0)
    x = np.arange(mx) * dx
    y = np.arange(my) * dy
    fac = 1.0 + alpha * x[None, :] * y[:, None] + beta * np.sin(4 * pi * x)[None, :] * np.sin(4 * pi * y)[:, None]
    fac_t = torch.from_numpy(fac.reshape(-1, 1)).cuda()
    acoef_t = torch.from_numpy(acoef).cuda()      # acoef[k, i]
    bcoef_t = torch.from_numpy(bcoef).cuda()

    def rates(c):                                  # c: [npts, ns]
        rxy = torch.zeros_like(c)
        for i in range(ns):                        # rxy(k) = rxy(k) + c(i)*acoef(k,i), i = 1..ns (:290-293)
            rxy = rxy + c[:, i:i + 1] * acoef_t[:, i][None, :]
        return c * (bcoef_t[None, :] * fac_t + rxy)   # :295-297

    return rates


def _run_jac(dk, ns, nsd, mx, my, u, rewt, cj, rates):
    import torch
    from daskr_b200.api import DeviceBuffer

    L = dk.lib()
    iw = np.zeros(64, dtype=np.int32)
    dk.setup_rbdpre(mx, my, ns, nsd, 0, iw)
    n, npts = u.size, mx * my
    calls = []

    def rblock(t, pmx, pmy, pns, pu, pr, user):
        assert C.cast(pmx, C.POINTER(C.c_int))[0] == mx and C.cast(pns, C.POINTER(C.c_int))[0] == ns
        L.dkb_sync()                               # the library's stream and torch's are different streams
        ut = _torch_view(pu, n).view(npts, ns)
        _torch_view(pr, n).view(npts, ns).copy_(rates(ut))
        torch.cuda.synchronize()
        calls.append(1)

    cb = RBLOCK_T(rblock)
    du, dw, dr1 = DeviceBuffer.from_host(u), DeviceBuffer.from_host(rewt), DeviceBuffer(8 * n)
    r0 = rates(torch.from_numpy(u).cuda().view(npts, ns)).contiguous()
    torch.cuda.synchronize()
    dbd, dip = DeviceBuffer(8 * ns * ns * npts), DeviceBuffer(4 * ns * npts)
    t, cjc, ierr = C.c_double(0.0), C.c_double(cj), C.c_int(-7)
    L.dkb_jac_rbdpre.restype = None
    L.dkb_jac_rbdpre.argtypes = [C.c_void_p] * 11
    L.dkb_jac_rbdpre(C.byref(t), du.ptr, C.c_void_p(r0.data_ptr()), C.cast(cb, C.c_void_p), None, dr1.ptr, dw.ptr, C.byref(cjc),
                     dbd.ptr, dip.ptr, C.byref(ierr))
    assert len(calls) == ns                        # one batched rblock call per perturbed species
    assert np.array_equal(du.to_host(np.float64, (n,)), u)   # u restored exactly (:229)
    return dbd, dip, ierr.value, r0.cpu().numpy().reshape(-1)


@pytest.mark.parametrize("np_,mx,my", [(1, 20, 20), (2, 9, 6), (10, 16, 12), (3, 33, 7)])
def test_user_rblock_food_web_rates_bit_exact(dk, oracle, np_, mx, my):
    from daskr_b200.api import DeviceBuffer

    ns = 2 * np_
    oracle.web_setup(np_, mx, my)
    c, _ = oracle.web_cinit(np_, mx, my)
    c = c * (1.0 + 0.01 * np.random.default_rng(1).standard_normal(c.size))
    rewt = 1.0 / (1e-5 * np.abs(c) + 1e-5)
    cj = 1.0 / 3.7e-4
    bd_o, ip_o, ierr_o = oracle.web_jac(c, rewt, cj, ns)
    dbd, dip, ierr, r0 = _run_jac(dk, ns, np_, mx, my, c, rewt, cj, _web_rates_torch(np_, mx, my))
    _, r0_o = oracle.web_res(c, np.zeros_like(c))
    assert np.array_equal(r0, r0_o), "the torch rate law must be the reference's arithmetic"
    assert ierr == ierr_o == 0
    npts = mx * my
    assert np.array_equal(dip.to_host(np.int32, (npts, ns)), ip_o), "pivots bit-exact"
    assert np.array_equal(dbd.to_host(np.float64, (npts, ns * ns)), bd_o), "LU factors bit-exact"
    b = np.random.default_rng(2).standard_normal(c.size)
    L = dk.lib()
    db = DeviceBuffer.from_host(b)
    L.dkb_psol_rbdpre(db.ptr, dbd.ptr, dip.ptr)
    assert L.dkb_sync() == 0
    x_o = oracle.psol_rbdpre(b, bd_o, ip_o)
    assert np.array_equal(db.to_host(np.float64, b.shape), x_o)
    # the same blocks in the streaming (tile-interleaved) layout
    dwp, diwp = DeviceBuffer(8 * L.dkb_rbd_lenwp()), DeviceBuffer(4 * L.dkb_rbd_leniwp())
    assert L.dkb_rbd_from_reference(dbd.ptr, dip.ptr, dwp.ptr, diwp.ptr) == 0
    db2 = DeviceBuffer.from_host(b)
    L.dkb_psol_rbdpre_tiles.restype = None
    L.dkb_psol_rbdpre_tiles(db2.ptr, dwp.ptr, diwp.ptr)
    assert L.dkb_sync() == 0
    assert np.array_equal(db2.to_host(np.float64, b.shape), x_o)


def test_user_rblock_other_reaction_term_and_singular_block(dk, oracle):
    """R_i = -k_i u_i^2 + s * u_(i+1 mod ns): analytic block Jacobian; with cj = 0 and nsd = 0, a point whose
    concentrations are all zero has a zero first column after the difference quotient -> dgefa info = 1 there."""
    import torch

    ns, mx, my = 3, 11, 5
    npts = mx * my
    k = torch.tensor([1.0, 2.0, 0.5], dtype=torch.float64, device="cuda")
    s = 0.25

    def rates(u):
        return (-k[None, :]) * u * u + s * torch.roll(u, -1, dims=1)

    rng = np.random.default_rng(7)
    u = 1.0 + rng.random(npts * ns)
    rewt = 1.0 / (1e-6 * np.abs(u) + 1e-6)
    cj = 50.0
    dbd, dip, ierr, _ = _run_jac(dk, ns, ns, mx, my, u, rewt, cj, rates)
    assert ierr == 0
    # LU factors of cj*I - dR/du (the FD quotient is exact to ~1e-7 for this quadratic term): check by solving
    L = dk.lib()
    from daskr_b200.api import DeviceBuffer

    b = rng.standard_normal(npts * ns)
    db = DeviceBuffer.from_host(b)
    L.dkb_psol_rbdpre(db.ptr, dbd.ptr, dip.ptr)
    assert L.dkb_sync() == 0
    x = db.to_host(np.float64, (npts, ns))
    up = u.reshape(npts, ns)
    kk = np.array([1.0, 2.0, 0.5])
    for p in (0, 17, npts - 1):
        jac = np.diag(-2 * kk * up[p]) + s * np.roll(np.eye(ns), 1, axis=1)
        a = cj * np.eye(ns) - jac
        assert np.allclose(a @ x[p], b.reshape(npts, ns)[p], rtol=1e-5, atol=1e-8)
    # a singular block: all-zero state at point 23, cj = 0, no differential species.  R = 0 there, the perturbed R of
    # column 1 is -k del^2 (row 1) and 0 elsewhere except the coupling s*del into row ns: the block is s*P - ..., regular.
    # Make it singular instead with s = 0: every off-diagonal entry vanishes and the diagonal is k*del ~ 1e-10 * ... != 0,
    # so use a reaction term that ignores species 1 altogether: column 1 is exactly zero.
    def rates2(w):
        r = torch.zeros_like(w)
        r[:, 1:] = -w[:, 1:]
        return r

    dbd, dip, ierr, _ = _run_jac(dk, ns, 0, mx, my, u, rewt, 0.0, rates2)
    assert ierr == 1                                  # dgefa: zero pivot in column 1 of the first failing block
