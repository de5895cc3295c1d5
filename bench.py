#!/usr/bin/env python
"""bench.py -- GMRES iterations/s of the DASKR Krylov path on the food-web DAE (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle restatement

A "step" is one BDF time step of DASKR (ddstp: predictor, 1-4 Newton iterations, each a
preconditioned GMRES solve = the hot path).  Workload at N=1: BASELINE configs[1], food web
ns=20 (np=10) on a 1024x1024 mesh, reaction-based block-diagonal preconditioner, rtol=atol=1e-5,
consistent initial values computed by DASKR (info(11)=1), tout=1.  For N>1 the mesh grows with
N along y (1024 x 1024*N, one 1024-row slab per GPU: weak scaling) and `value` is reported in
1024x1024-mesh equivalents (NLI/s x N) so that it is the whole-job aggregate.

value  : dNLI/dt over K steps after W warm-up steps, state resident in HBM (one dkb_daskr call,
         device y/yprime), timed with CUDA events on the library's stream, max over ranks.
e2e    : the same K-step window through dkb_daskr with pinned HOST y/yprime in
         intermediate-output mode (info(3)=1): one call per step, y and yprime copied
         device->host at every return.  By DASKR's convention y/yprime are inputs only on the
         initial call, so the per-step host->device payload is the scalar headers; the one-off
         upload of the initial state is timed separately (h2d_bytes_initial / initial_call_ms).
roofline: the fused datv kernel (residual + block solve, K3), algorithmic bytes
         N*[8*(6+ns)+4] per launch / mean launch time from CUDA events recorded
         around every launch inside the timed region.
cpu_baseline: the oracle (C++ restatement of the reference's Fortran, 1 core: the reference is
         serial) on a bounded sample of the workload, rank 0, N=1 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NP_SPECIES = 10          # ns = 20
MX = 1024
MY_PER_GPU = 1024
RTOL = ATOL = 1e-5
TOUT = 1.0


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                clk = float(f[0])
                smax = float(f[1])
            except ValueError:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # window shorter than the sampling period: take every sample
            for ts, line in self.rows:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[0]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------- CPU arm (oracle)
def cpu_sample(steps, warmup, mesh=None):
    """Time the oracle (CPU restatement) on a bounded sample: the same food-web ns=20 problem on a
    `mesh` x `mesh` sub-mesh, W+K steps, NLI/s over the last K steps scaled to the 1024^2 mesh by
    the unknown ratio (every sweep of the path is O(N))."""
    from tests import oracle_py

    orc = oracle_py.load()
    total = steps + warmup
    if mesh is None:
        mesh = 256 if total <= 70 else 128
    import ctypes as C

    logbuf = np.zeros((total + 8, 3))
    orc.L.o_set_steplog.argtypes = [C.c_void_p, C.c_long]
    orc.L.o_set_steplog(logbuf.ctypes.data_as(C.c_void_p), total + 8)
    t0 = time.time()
    r = orc.run_web(NP_SPECIES, mesh, mesh, [TOUT], jpre=1, rtol=RTOL, atol=ATOL, max_steps=total)
    wall = time.time() - t0
    n = int(orc.L.o_steplog_count())
    orc.L.o_set_steplog(None, 0)
    if n < total or r["idid"] < 0:
        raise RuntimeError("oracle sample run ended early: idid=%d steps=%d" % (r["idid"], n))
    w = max(warmup, 1)
    dt = logbuf[total - 1, 0] - logbuf[w - 1, 0]
    dnli = logbuf[total - 1, 1] - logbuf[w - 1, 1]
    frac = (mesh * mesh) / float(MX * MY_PER_GPU)
    rate_sample = dnli / dt
    return {
        "value": rate_sample * frac, "unit": "it/s", "cores": 1, "kind": "port",
        "sample": "oracle (C++ restatement of the Fortran reference; no Fortran compiler in the image), food web "
                  "ns=20 on a %dx%d sub-mesh (%.4g of the unknowns), steps %d..%d of the same run: %.0f GMRES its in "
                  "%.2f s = %.3f it/s at sample size, scaled by the unknown ratio" % (mesh, mesh, frac, w + 1, total, dnli, dt, rate_sample),
        "nproc": os.cpu_count(), "wall_s": wall, "nli": float(dnli), "ms_per_step": 1e3 * dt / max(total - w, 1),
    }


def cpu_sample_native(steps, warmup, mesh=None):
    """Courtesy number: the same oracle sources rebuilt on THIS host with -O3 -march=native
    (FMA contraction allowed, so not the parity arithmetic) and timed on the same sample in a child process."""
    import tempfile

    odir = os.path.join(ROOT, "oracle")
    srcs = [os.path.join(odir, f) for f in ("blas_linpack.cpp", "krylov.cpp", "rbdpre_problems.cpp", "daskr_driver.cpp")]
    with tempfile.TemporaryDirectory() as tmp:
        lib = os.path.join(tmp, "liboracle.so")
        subprocess.run(["g++", "-O3", "-march=native", "-fPIC", "-std=c++17", "-shared", "-o", lib] + srcs + ["-lm"], check=True)
        code = ("import sys, json; sys.path.insert(0, %r); import bench; from tests import oracle_py; oracle_py.LIB = %r; "
                "oracle_py.ORACLE_DIR = %r; r = bench.cpu_sample(%d, %d, %r); print(json.dumps({k: r[k] for k in ('value', 'nli', 'wall_s')}))"
                % (ROOT, lib, tmp, steps, warmup, mesh))
        out = subprocess.run([sys.executable, "-c", code], check=True, stdout=subprocess.PIPE, text=True).stdout
    return json.loads(out.strip().splitlines()[-1])


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cb = cpu_sample(args.steps, args.warmup)
    if args.cpu_native:
        try:
            cb["sample"] += "; courtesy rebuild with -O3 -march=native on this host: %.3f it/s" % cpu_sample_native(args.steps, args.warmup)["value"]
        except Exception as ex:
            cb["sample"] += "; courtesy -O3 -march=native run failed: %r" % (ex,)
    line = {
        "impl": "reference", "metric": "gmres_iters_per_s", "value": cb["value"], "unit": "it/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "food-web DAE ns=20 (np=10), 1024x1024 mesh, drbdpre block-diagonal preconditioner, "
                               "rtol=atol=1e-5 (BASELINE configs[1]); CPU arm runs a bounded sub-mesh sample"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "nproc")},
        "e2e": {"value": cb["value"], "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    import daskr_b200 as dk
    from daskr_b200.api import DeviceBuffer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log("warning: --gpus %d but WORLD_SIZE=%d; using WORLD_SIZE" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- daskr_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dk.init(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the one JSON line: NCCL prints its version banner to stdout at any debug level
        os.environ.pop("NCCL_DEBUG", None)
        if "DKB_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["DKB_NCCL_DEBUG"]
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)   # anything the communicator set-up prints goes to stderr
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dk.comm_init_from_torch(rank, world)
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    K, W = args.steps, args.warmup
    ns = 2 * NP_SPECIES
    mx, my = args.mx, args.my_per_gpu * world
    probe = np.zeros(64, dtype=np.int32)
    dk.setup_rbdpre(mx, my, ns, NP_SPECIES, 0, probe)
    jy0, myloc = dk.slab()
    neq = ns * mx * myloc
    neq_g = ns * mx * my
    iwork = np.zeros(40 + neq, dtype=np.int32)
    dk.setup_rbdpre(mx, my, ns, NP_SPECIES, 40, iwork)
    ay = float(world) if args.ay is None else float(args.ay)
    dk.web_setpar(NP_SPECIES, mx, my, ay=ay)   # default: N unit squares stacked in y, same mesh spacing per GPU
    stream = torch.cuda.ExternalStream(dk.lib().dkb_stream())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- consistent initial values (untimed set-up; host arrays) ----
    info = np.zeros(20, dtype=np.int32)
    info[10] = 1; info[11] = 1; info[13] = 1; info[14] = 1; info[15] = 1
    rwork = np.zeros(60)
    rpar = np.zeros(1)
    ipar = np.array([args.jpre, 0], dtype=np.int32)
    # pinned host state
    c_t = torch.empty(neq, dtype=torch.float64).pin_memory()
    cd_t = torch.empty(neq, dtype=torch.float64).pin_memory()
    c, cdot = c_t.numpy(), cd_t.numpy()
    c0, cdot0 = dk.web_cinit(neq, 1e5 * NP_SPECIES)   # predator guess ee*10*np (example: 1e5 at np=1)
    c[:] = c0
    cdot[:] = cdot0
    del c0, cdot0
    dk.set_option(dk.OPT_DEVICE_IO, 0)
    t0 = time.time()
    t, idid = dk.daskr(dk.WEB_RES, neq, 0.0, c, cdot, TOUT, info, RTOL, ATOL, rwork, iwork, rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)
    if idid != 4:
        raise SystemExit("initial-condition calculation failed: idid=%d" % idid)
    log("[rank %d] IC calculation %.2fs, counters nre=%d nje=%d nni=%d nli=%d" % (rank, time.time() - t0, iwork[11], iwork[12], iwork[18], iwork[19]))
    c_ic, cdot_ic = c.copy(), cdot.copy()
    info[10] = 0

    def events():
        return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # =========== (1) device-resident: W warm-up steps, K timed steps, one call each ===========
    d_y, d_yp = DeviceBuffer.from_host(c_ic), DeviceBuffer.from_host(cdot_ic)
    dk.set_option(dk.OPT_DEVICE_IO, 1)
    info[0] = 0
    info[2] = 0
    tcur = 0.0

    def advance(nsteps, y, yp):
        """one dkb_daskr call that takes exactly nsteps BDF steps (per-call step limit -> idid=-1)"""
        nonlocal tcur
        dk.set_option(dk.OPT_MAX_STEPS, nsteps)
        nst0 = int(iwork[10]) if info[0] != 0 else 0
        tcur, idid = dk.daskr(dk.WEB_RES, neq, tcur, y, yp, TOUT, info, RTOL, ATOL, rwork, iwork, rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)
        if idid == -1:
            info[0] = 1   # "too many steps on this call": continue (DASKR doc, idid=-1)
        elif idid < 0:
            raise SystemExit("daskr failed in the benchmark window: idid=%d at t=%g" % (idid, tcur))
        return int(iwork[10]) - nst0

    if W > 0:
        advance(W, d_y.ptr.value, d_yp.ptr.value)
    barrier()
    dk.prof_reset()
    dk.prof_enable(True)
    cnt0 = iwork[:40].copy()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    e0, e1 = events()
    barrier()
    w0 = time.time()
    e0.record(stream)
    taken = advance(K, d_y.ptr.value, d_yp.ptr.value)
    e1.record(stream)
    barrier()
    w1 = time.time()
    clocks = sampler.stop(w0, w1)
    ms = e0.elapsed_time(e1)
    prof = dk.prof_get()
    launches = dk.launch_count()
    dk.prof_enable(False)
    cnt1 = iwork[:40].copy()
    if taken != K:
        raise SystemExit("timed window took %d steps instead of %d" % (taken, K))
    tm = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms = float(tm.item())
    dnli = int(cnt1[19] - cnt0[19])
    dcount = {k: int(cnt1[i] - cnt0[i]) for k, i in (("nst", 10), ("nre", 11), ("nje", 12), ("netf", 13), ("ncfn", 14), ("ncfl", 15), ("nni", 18), ("nli", 19), ("nps", 20))}
    scale_units = (neq_g / float(ns * MX * MY_PER_GPU))   # 1024x1024-mesh equivalents
    value = dnli / (ms * 1e-3) * scale_units
    t_end_window = tcur
    log("[rank %d] device-resident: %d steps in %.1f ms, dNLI=%d, counters %s, t=%.3e h=%.2e" % (rank, K, ms, dnli, dcount, tcur, rwork[6]))

    # roofline of the dominant kernel (fused datv)
    dat = prof.get("datv", {"ms": 0.0, "count": 0})
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    # Dominant kernel: tpp_psol_kernel<20> (psol_rbdpre fused with "- savr" and "* wght").  Algorithmic
    # bytes per launch: the LU block (8*ns per unknown) and pivots (4) once, plus the vectors it must
    # touch: in, wt, out (+ savr inside datv).  It runs once per Krylov iteration inside datv and once
    # per dspigm call as the prologue psol (no savr).
    tpp = prof.get("tpp_psol_kernel", {"ms": 0.0, "count": 0})
    n_datv = dat["count"]
    n_psol0 = max(tpp["count"] - n_datv, 0)
    b_datv = neq * (8 * ns + 4 + 8 * 4)
    b_psol0 = neq * (8 * ns + 4 + 8 * 3)
    roof = {"bound": "hbm", "kernel": "tpp_psol_kernel<20> (psol_rbdpre block solve + datv's '- savr' / '* wght')",
            "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": b_datv, "launches": tpp["count"]}
    if tpp["count"] > 0 and tpp["ms"] > 0:
        ach = (n_datv * b_datv + n_psol0 * b_psol0) / (tpp["ms"] * 1e-3) / 1e9
        roof["achieved"] = ach
        roof["frac"] = ach / peak
        roof["avg_launch_ms"] = tpp["ms"] / tpp["count"]
        roof["share_of_step_time"] = tpp["ms"] / ms
    # the whole Jacobian-vector product (residual kernel + block solve) against the bytes a single
    # fully fused kernel would have to move, N*[8*(6+ns)+4] (SURVEY.md 8d "phase A")
    if dat["count"] > 0 and dat["ms"] > 0:
        b_phase = neq * (8 * (6 + ns) + 4)
        ach = b_phase / (dat["ms"] / dat["count"] * 1e-3) / 1e9
        roof["datv_phase"] = {"kernels": "web_restile_kernel<20,true> + tpp_psol_kernel<20>", "algorithmic_bytes": b_phase,
                              "avg_ms": dat["ms"] / dat["count"], "achieved": ach, "frac": ach / peak,
                              "share_of_step_time": dat["ms"] / ms}
    key = "tpp_psol_ns20_%dx%d" % (mx, myloc)
    # With the single-kernel datv (default when mx % 32 == 0) the dominant kernel is
    # web_datv_single_kernel<20>: residual at the perturbed state + block solve, N*[8*(6+ns)+4] bytes.
    one = prof.get("datv_single_kernel", {"ms": 0.0, "count": 0})
    if one["count"] > 0 and one["ms"] > 0:
        b_one = neq * (8 * (6 + ns) + 4)
        ach = b_one / (one["ms"] / one["count"] * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "web_datv_single_kernel<20> (fused datv: perturbed-state residual + psol_rbdpre block solve)",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": b_one, "launches": one["count"], "avg_launch_ms": one["ms"] / one["count"],
                "share_of_step_time": one["ms"] / ms}
        if tpp["count"] > 0:
            roof["psol_prologue_kernel"] = {"kernel": "tpp_psol_kernel<20>", "avg_launch_ms": tpp["ms"] / tpp["count"],
                                            "achieved": b_psol0 / (tpp["ms"] / tpp["count"] * 1e-3) / 1e9}
        key = "web_datv_single_ns20_%dx%d" % (mx, myloc)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        if key in tr:
            roof["traffic"] = tr[key]["dram_bytes_per_launch"]
            roof["traffic_source"] = tr[key].get("source")
    except Exception:
        pass
    # K5 (jac_rbdpre: FD block + LU) is the one compute-shaped kernel of the path: report FP64 rate.
    # flops per mesh point: ns rblock evaluations of 2*ns^2 + 3*ns, one R0, LINPACK dgefa sum_m (m + 2 m^2)
    jc = prof.get("jac", {"ms": 0.0, "count": 0})
    if jc["count"] > 0 and jc["ms"] > 0:
        fl_pt = (ns + 1) * (2 * ns * ns + 3 * ns) + sum(m + 2 * m * m for m in range(1, ns))
        roof["jac_kernel"] = {"kernel": "web_jac2_kernel<20,16> (FD block Jacobian + LU, incl. the first-failure flag readback)",
                              "calls": jc["count"], "avg_ms": jc["ms"] / jc["count"], "flops_per_point": fl_pt,
                              "achieved_gflops": fl_pt * (neq // ns) / (jc["ms"] / jc["count"] * 1e-3) / 1e9,
                              "fp64_peak_gflops": {"dfma": 33600.0, "dmul_dadd_pairs_no_fma": 18400.0,
                                                   "source": "tools/fp64_peak.cu on this GPU type (profiles/README.md)"},
                              "share_of_step_time": jc["ms"] / ms}
    phase_ms = {k: round(v["ms"], 3) for k, v in prof.items()}

    # =========== (2) e2e: same window through the host-buffer API, one call per step ===========
    dk.set_option(dk.OPT_DEVICE_IO, 0)
    info[0] = 0
    info[2] = 1   # info(3)=1: return after every step (y, yprime copied to the host each time)
    tcur = 0.0
    c[:] = c_ic
    cdot[:] = cdot_ic
    dk.set_option(dk.OPT_MAX_STEPS, 500)
    barrier()
    ei0, ei1 = events()
    ei0.record(stream)
    tcur, idid = dk.daskr(dk.WEB_RES, neq, tcur, c, cdot, TOUT, info, RTOL, ATOL, rwork, iwork, rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)
    ei1.record(stream)
    torch.cuda.synchronize()
    initial_call_ms = ei0.elapsed_time(ei1)   # H2D of y,yprime + first step + D2H
    if idid < 0:
        raise SystemExit("e2e initial call failed idid=%d" % idid)
    for _ in range(max(W - 1, 0)):
        tcur, idid = dk.daskr(dk.WEB_RES, neq, tcur, c, cdot, TOUT, info, RTOL, ATOL, rwork, iwork, rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)
        if idid < 0:
            raise SystemExit("e2e warm-up failed idid=%d" % idid)
    barrier()
    nli0 = int(iwork[19])
    e0, e1 = events()
    e0.record(stream)
    for _ in range(K):
        tcur, idid = dk.daskr(dk.WEB_RES, neq, tcur, c, cdot, TOUT, info, RTOL, ATOL, rwork, iwork, rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)
        if idid < 0:
            raise SystemExit("e2e window failed idid=%d" % idid)
    e1.record(stream)
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    tm = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms_e2e = float(tm.item())
    dnli_e2e = int(iwork[19]) - nli0
    e2e_value = dnli_e2e / (ms_e2e * 1e-3) * scale_units
    hdr_bytes = 60 * 8 + 40 * 4 + 20 * 4 + 5 * 8
    e2e = {"value": e2e_value, "unit": "it/s", "h2d_bytes_per_step": hdr_bytes, "d2h_bytes_per_step": 16 * neq,
           "h2d_bytes_initial": 16 * neq, "initial_call_ms": initial_call_ms, "ms_per_step": ms_e2e / K, "nli": dnli_e2e,
           "mode": "pinned host y/yprime, info(3)=1: one dkb_daskr call per step, y+yprime D2H at every return; "
                   "y/yprime are inputs only on the initial call (DASKR convention), so per-step H2D is the scalar headers"}
    log("[rank %d] e2e: %d steps in %.1f ms, dNLI=%d" % (rank, K, ms_e2e, dnli_e2e))

    # =========== (3) CPU baseline (rank 0, N=1 only) ===========
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cb = cpu_sample(min(K, 20), min(W, 30), mesh=args.cpu_mesh)
            cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "nproc")}
        except Exception as ex:   # the CPU leg must never take the GPU number down with it
            cpu = {"value": None, "unit": "it/s", "cores": 1, "kind": "port", "sample": "failed: %r" % (ex,)}

    if rank == 0:
        line = {
            "metric": "gmres_iters_per_s", "value": value, "unit": "it/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak" if args.ay is None else "strong",
            "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": "food-web DAE ns=20 (np=10), %dx%d mesh (%d rows per GPU), drbdpre block-diagonal "
                            "preconditioner, rtol=atol=1e-5, IC calculation on, tout=1 (BASELINE configs[%d]%s)"
                            % (mx, my, myloc, 3 if (mx == 4096 and my == 4096) else 1,
                               "" if (world == 1 and mx == MX) else "; value in 1024x1024-mesh equivalents"
                               + (", mesh grown along y for weak scaling" if args.ay is None and world > 1 else "")),
                "neq_global": neq_g, "domain": "[0,1]x[0,%g]" % ay, "jpre": args.jpre, "gmres": "MAXL=5 KMP=5 NRMAX=5 EPLI=0.05 (DASKR defaults)",
                "l2": "inputs larger than L2: %.0f MB per length-N vector, %.2f GB of LU blocks streamed per Krylov iteration" % (8 * neq / 1e6, 8.0 * ns * neq / 1e9),
                "window": "steps %d..%d after the IC calculation, t_end=%.3e" % (W + 1, W + K, t_end_window),
                "counters": dcount,
            },
            "nli_per_s": dnli / (ms * 1e-3), "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roof, "phase_ms": phase_ms,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=30)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mx", type=int, default=MX)
    ap.add_argument("--my-per-gpu", type=int, default=MY_PER_GPU)
    ap.add_argument("--ay", type=float, default=None,
                    help="domain height (default N: N unit squares stacked in y = weak scaling; 1 = the unit square split "
                         "into slabs, e.g. --mx 4096 --my-per-gpu 512 --ay 1 at N=8 is BASELINE configs[3])")
    ap.add_argument("--jpre", type=int, default=1, help="food-web preconditioner: 1 = drbdpre (configs[1]); 3 = as shipped")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-native", action="store_true",
                    help="--impl reference only: also time a -O3 -march=native rebuild of the oracle (courtesy number)")
    ap.add_argument("--cpu-mesh", type=int, default=None)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("note: W >= 3 warm-up steps are required for a valid number; got %d" % args.warmup)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
