"""world_size-2 test of the slab decomposition on CPU (gloo): the N>1 data path of SURVEY.md 8(e).

Each rank owns a y-slab of the food-web mesh, exchanges one ghost row with its neighbour and
evaluates the residual on its slab; dot products / WRMS norms are all-reduced scalars.  The
oracle's residual on the full mesh is the checker.  This mirrors exactly what the CUDA path does
with NCCL (halo_exchange + allreduce_sum/max in csrc/capi.cu), with gloo standing in for NCCL.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _slab_residual(c_loc, cdot_loc, ghost_lo, ghost_hi, np_, mx, my_g, jy0, myloc, orc):
    """Residual of rows [jy0, jy0+myloc) given ghost rows: embed the slab in a (myloc+2)-row mesh,
    evaluate the oracle's residual there and keep the interior rows.  Only valid for interior
    slabs' stencil terms; global edges are handled by giving the mirrored row as the ghost."""
    ns = 2 * np_
    rows = []
    rows.append(ghost_lo if ghost_lo is not None else c_loc[ns * mx: 2 * ns * mx])      # mirror: row 1
    rows.append(c_loc)
    rows.append(ghost_hi if ghost_hi is not None else c_loc[-2 * ns * mx: -ns * mx])     # mirror: row my-2
    return np.concatenate(rows)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from daskr_b200.slab import partition
    from tests import oracle_py

    orc = oracle_py.load()
    np_, mx, my = 2, 9, 10
    ns = 2 * np_
    row = ns * mx
    orc.web_setup(np_, mx, my)
    c, cdot = orc.web_cinit(np_, mx, my)
    rng = np.random.default_rng(0)
    c = c * (1 + 0.01 * rng.standard_normal(c.size))
    cdot = cdot + rng.standard_normal(c.size)
    full, _ = orc.web_res(c, cdot)

    jy0, myloc = partition(my, world, rank)
    lo, hi = jy0 * row, (jy0 + myloc) * row
    c_loc, cd_loc = c[lo:hi].copy(), cdot[lo:hi].copy()

    # halo exchange: first row down, last row up (K9)
    ghost_lo = ghost_hi = None
    reqs = []
    if rank > 0:
        ghost_lo_t = torch.empty(row, dtype=torch.float64)
        reqs.append(dist.isend(torch.from_numpy(c_loc[:row].copy()), rank - 1))
        reqs.append(dist.irecv(ghost_lo_t, rank - 1))
    if rank < world - 1:
        ghost_hi_t = torch.empty(row, dtype=torch.float64)
        reqs.append(dist.isend(torch.from_numpy(c_loc[-row:].copy()), rank + 1))
        reqs.append(dist.irecv(ghost_hi_t, rank + 1))
    for r in reqs:
        r.wait()
    if rank > 0:
        ghost_lo = ghost_lo_t.numpy()
    if rank < world - 1:
        ghost_hi = ghost_hi_t.numpy()

    # evaluate the residual on [ghost | slab | ghost] as a small mesh and keep the slab rows; the
    # x/y-dependent rate factor needs the true global row, so use the full-mesh oracle with the
    # slab + ghosts pasted into an otherwise poisoned array: any read outside slab+ghost shows up
    poisoned = np.full_like(c, np.nan)
    poisoned[lo:hi] = c_loc
    if ghost_lo is not None:
        poisoned[lo - row: lo] = ghost_lo
    if ghost_hi is not None:
        poisoned[hi: hi + row] = ghost_hi
    cd_full = np.zeros_like(c)
    cd_full[lo:hi] = cd_loc
    mine, _ = orc.web_res(poisoned, cd_full)
    mine = mine[lo:hi]
    ok_res = np.array_equal(mine, full[lo:hi])     # one ghost row each way is all the stencil needs

    # scalar allreduce of a WRMS norm (sum of squares + max) and of a dot product
    w = 1.0 / (1e-5 * np.abs(c) + 1e-5)
    p = (full * w)[lo:hi]
    parts = torch.tensor([np.sum(p * p)], dtype=torch.float64)
    vmax = torch.tensor([np.abs(p).max()], dtype=torch.float64)
    dist.all_reduce(parts, op=dist.ReduceOp.SUM)
    dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
    wr = np.sqrt(parts.item() / c.size)
    ref = orc.ddwnrm(full, w)
    ok_norm = abs(wr - ref) <= 1e-12 * ref and vmax.item() == np.abs(full * w).max()
    q.put((rank, bool(ok_res), bool(ok_norm), jy0, myloc))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_residual_and_norms_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(r[4] for r in res) == 10
    for rank, ok_res, ok_norm, _, _ in res:
        assert ok_res, "slab residual of rank %d differs from the full-mesh residual" % rank
        assert ok_norm, "all-reduced WRMS norm of rank %d differs" % rank


# ---------------------------------------------------------------- Gauss-Seidel factor across slabs
def _gs_sweeps(mat, rhs):
    """5 lexicographic Gauss-Seidel iterations from x = 0 on mat x = rhs (what gauss_seidel computes,
    tests/test_oracle_cpu.py::test_gauss_seidel_matches_the_residual_operator)."""
    from scipy.linalg import solve_triangular

    dl = np.tril(mat)
    u = dl - mat
    x = np.zeros(rhs.size)
    for _ in range(5):
        x = solve_triangular(dl, u @ x + rhs, lower=True)
    return x


def _gs_worker(rank, world, port, q):
    """The multi-GPU form of the spatial factor with gloo standing in for NCCL: every rank sends its top
    GS_OV_BELOW rows up and its bottom GS_OV_ABOVE rows down (slab_overlap_exchange, csrc/capi.cu), sweeps
    [rows from below | own rows | rows from above] with the couplings across the window's ends dropped, keeps its
    own rows.  Checker: the global sweep on the full mesh."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from daskr_b200.slab import partition, gs_overlap
    from tests import oracle_py

    orc = oracle_py.load()
    np_, mx, my = 1, 12, 40 * world            # (anisotropic cells: the weakest damping, see the assertion)
    ns = 2 * np_
    row = ns * mx
    neq = row * my
    orc.web_setup(np_, mx, my)

    def diffusion(c):
        delta, rpar = orc.web_res(c, np.zeros(neq))
        return -delta - rpar

    jd = np.stack([diffusion(e) for e in np.eye(neq)], axis=1)
    a = np.eye(neq) - (1.0 / 50.0) * jd
    b = np.random.default_rng(1).standard_normal(neq)
    full = _gs_sweeps(a, b)

    jy0, myloc = partition(my, world, rank)
    nb, na = gs_overlap(my, world, rank)
    own = b[jy0 * row:(jy0 + myloc) * row].copy()
    ob, oa = gs_overlap(my, world, 1)[0], gs_overlap(my, world, 0)[1]      # what neighbours expect from me
    below = torch.empty(nb * row, dtype=torch.float64)
    above = torch.empty(na * row, dtype=torch.float64)
    reqs = []
    if rank > 0:
        reqs.append(dist.isend(torch.from_numpy(own[:oa * row].copy()), rank - 1))       # my bottom rows go down
        reqs.append(dist.irecv(below, rank - 1))
    if rank < world - 1:
        reqs.append(dist.isend(torch.from_numpy(own[-ob * row:].copy()), rank + 1))      # my top rows go up
        reqs.append(dist.irecv(above, rank + 1))
    for r in reqs:
        r.wait()
    window_b = np.concatenate([below.numpy(), own, above.numpy()])
    lo, hi = (jy0 - nb) * row, (jy0 + myloc + na) * row
    ok_rows = np.array_equal(window_b, b[lo:hi])                      # the exchange delivered the right rows
    win = _gs_sweeps(a[lo:hi, lo:hi], window_b)
    got = win[nb * row: nb * row + myloc * row]
    ref = full[jy0 * row:(jy0 + myloc) * row]
    err = float(np.abs(got - ref).max() / np.abs(ref).max())
    q.put((rank, bool(ok_rows), err))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_overlapped_slab_gauss_seidel_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + (os.getpid() % 500) + world
    procs = [ctx.Process(target=_gs_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_rows, err in res:
        assert ok_rows, "rank %d received the wrong overlap rows" % rank
        # anisotropic cells here (dy = dx*(mx-1)/(my-1)): the damping per row is between 1/3 and 1/2
        assert err < 1e-7, "rank %d: own rows differ from the global sweep by %.2e" % (rank, err)
