"""Multi-GPU parity check (run under torchrun on a GPU box; tests/test_gpu_multi.py launches it from pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/mgpu_check.py

Every rank owns a y-slab, halo rows travel by NCCL send/recv, dots and norms by NCCL allreduce.
Rank 0 gathers the slabs and compares with the oracle's full-mesh run (within 10x tolerance in
the WRMS norm; counters side by side).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
    import daskr_b200 as dk

    dk.init(local)
    dk.comm_init_from_torch(rank, world)
    ok = True
    # jpre = 3: the spatial Gauss-Seidel factor sweeps overlapped slabs on several GPUs (gs.cu), so those cases agree with
    # the full-mesh CPU run to ~1e-8 x tolerance, not bit for bit
    cases = [("web", 1, 24, 26, True, 1, 0), ("web", 1, 24, 26, False, 1, 0), ("web", 10, 16, 18, True, 1, 0),
             ("heat", 0, 30, 0, True, 1, 0), ("web", 1, 40, 44, True, 3, 0), ("web", 10, 36, 40, True, 3, 0),
             ("web", 1, 40, 44, True, 3, 5), ("web", 2, 30, 33, True, 1, 4),   # last column: block grouping, ng x ng groups
             ("web", 2, 32, 36, True, 3, 0), ("web", 1, 64, 30, True, 1, 0),   # mx % 32 == 0: single-kernel datv / residual
             ("web", 1, 96, 40, True, 3, -32), ("web", 2, 160, 48, True, 3, -64)]  # ng < 0: Gauss-Seidel column panels of -ng columns
    for kind, np_, mx, my, fused, jpre, ng in cases:
        os.environ.pop("DKB_GS_PANEL_COLS", None)
        if ng < 0:
            os.environ["DKB_GS_PANEL_COLS"] = str(-ng)
            ng = 0
        if kind == "web":
            touts = [1e-8, 1e-6, 1e-4, 1e-3, 1e-2]
            g = dk.solve_web(np_, mx, my, touts, fused=fused, jpre=jpre, nxg=ng, nyg=ng)
            rtol, atol = 1e-5, 1e-5
        else:
            touts = [0.01 * 2 ** n for n in range(6)]
            g = dk.solve_heat(mx, touts, fused=fused)
            rtol, atol = 0.0, 1e-5
        ys = g["y"]
        gathered = [None] * world
        dist.all_gather_object(gathered, ys)
        if rank == 0:
            from tests import oracle_py

            orc = oracle_py.load()
            o = orc.run_web(np_, mx, my, touts, jpre=jpre, nxg=ng, nyg=ng) if kind == "web" else orc.run_heat(mx, touts)
            full = np.concatenate(gathered, axis=1)
            assert full.shape == o["y"].shape, (full.shape, o["y"].shape)
            w = 1.0 / (rtol * np.abs(o["y"]) + atol)
            err = np.sqrt(np.mean(((full - o["y"]) * w) ** 2, axis=1))
            cg, co = g["counters"][-1], o["counters"][-1]
            good = g["idid"] == o["idid"] == 3 and err.max() <= 10.0
            ok = ok and good
            print("%s np=%d %dx%d fused=%s jpre=%d groups=%d world=%d: max wrms(diff)/tol=%.3e  NST %d/%d NNI %d/%d NLI %d/%d NJE %d/%d -> %s"
                  % (kind, np_, mx, my, fused, jpre, ng, world, err.max(), cg[10], co[10], cg[18], co[18], cg[19], co[19], cg[12], co[12],
                     "OK" if good else "FAIL"), flush=True)
    # ---- rootfinding across slabs: the root function is an all-reduced device sum (example_web.f90:577-617) ----
    touts = [1e-6, 1e-4, 1e-2, 5e-2]
    g = dk.solve_web(1, 20, 22, touts, roots=True)
    if rank == 0:
        from tests import oracle_py
        from tests.test_gpu_roots import oracle_web_roots

        o = oracle_web_roots(oracle_py.load(), 1, 20, 22, touts)
        good = len(g["roots"]) == len(o["roots"]) == 1 and abs(g["roots"][0][0] - o["roots"][0][0]) <= 1e-9 * o["roots"][0][0] \
            and g["roots"][0][1][0] == o["roots"][0][1][0] and g["roots"][0][2][10] == o["roots"][0][2][10]
        ok = ok and good
        print("roots world=%d: t_root %.12e / %.12e jroot %d/%d NST %d/%d -> %s" % (
            world, g["roots"][0][0], o["roots"][0][0], g["roots"][0][1][0], o["roots"][0][1][0], g["roots"][0][2][10],
            o["roots"][0][2][10], "OK" if good else "FAIL"), flush=True)

    # ---- checkpoint / continuation with one state buffer per rank ----
    def run(break_after):
        probe = np.zeros(64, dtype=np.int32)
        dk.setup_rbdpre(24, 26, 2, 1, 0, probe)
        _, myloc = dk.slab()
        neq = 2 * 24 * myloc
        iwork = np.zeros(40 + neq, dtype=np.int32)
        dk.setup_rbdpre(24, 26, 2, 1, 40, iwork)
        dk.web_setpar(1, 24, 26)
        info = np.zeros(20, dtype=np.int32)
        info[[10, 11, 13, 14, 15]] = 1
        rwork, rpar, ipar = np.zeros(60), np.zeros(1), np.array([3, 0], dtype=np.int32)
        c, cdot = dk.web_cinit(neq)
        t, idid = dk.daskr(dk.WEB_RES, neq, 0.0, c, cdot, 1e-6, info, 1e-5, 1e-5, rwork, iwork, rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)
        info[10] = 0
        outs = []
        for k, tout in enumerate([1e-6, 1e-4, 1e-3, 1e-2]):
            if k == break_after:
                state = dk.state_export()
                dk.web_setpar(1, 24, 26)          # (the communicator lives on: same process, state dropped and restored)
                dk.state_import(state)
            t, idid = dk.daskr(dk.WEB_RES, neq, t, c, cdot, tout, info, 1e-5, 1e-5, rwork, iwork, rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)
            outs.append((c.copy(), iwork[:40].copy()))
        return outs

    a, b = run(None), run(2)
    same = all(np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]) for x, y in zip(a, b))
    st = torch.tensor([1 if same else 0])
    dist.all_reduce(st, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("state export/import world=%d: continuation bitwise %s" % (world, "OK" if st.item() == 1 else "FAIL"), flush=True)
        ok = ok and st.item() == 1
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
