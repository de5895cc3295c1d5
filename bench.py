#!/usr/bin/env python
"""bench.py -- GMRES iterations/s of the DASKR Krylov path on the food-web DAE (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle restatement

A "step" is one BDF time step of DASKR (ddstp: predictor, 1-4 Newton iterations, each a preconditioned
GMRES solve = the hot path).  Default workload = BASELINE configs[3], the configuration the north-star
target is quoted on: food web ns=20 (np=10) on the 4096x4096 mesh of the unit square, rtol=atol=1e-5,
consistent initial values computed by DASKR (info(11)=1), tout=1, preconditioner as shipped by the
reference example (jpre=3: Gauss-Seidel factor x reaction block factor; with the reaction factor alone the
IC calculation does not converge on this mesh -- tests/test_gpu_fullsize.py keeps that on record).  The mesh
rows are split over the N GPUs (4096/N rows each): STRONG scaling, `value` is the plain dNLI/dt of the job.

value   : dNLI/dt over K steps after W warm-up steps, state resident in HBM (one dkb_daskr call, device
          y/yprime), timed with CUDA events on the library's stream, max over ranks.
drbdpre : a second device-resident window on the same run with ipar(1) switched to 1 (the block-diagonal
          preconditioner alone, what BASELINE words as "drbdpre"): its Krylov iteration is ONE kernel + MGS.
e2e     : the first window again through dkb_daskr with pinned HOST y/yprime in intermediate-output mode
          (info(3)=1): one call per step, y and yprime copied device->host at every return (DASKR makes them
          inputs only on the initial call, whose upload is timed separately).
roofline: the dominant kernel of the timed window (CUDA events around every launch), plus `whole_iteration`:
          SURVEY 8(d)'s algorithmic bytes B_it(ll) = N[8(6+ns)+4] + 8N(4ll+3) summed over the measured ll
          histogram / the event-timed Arnoldi iterations, against the measured and the nominal HBM peak.
digest  : mean(c1), wrms(y), t, h and all counters at the end of the timed window, so that the runs at
          N = 2, 4, 8 can be checked against N = 1 from the JSON lines alone.
cpu_baseline: the oracle (C++ restatement of the reference's Fortran, 1 core: the reference is serial) on a
          bounded sub-mesh sample of the same algorithm, rank 0, N=1 only.

Other workloads: --workload web1024 (configs[1]: 1024^2, drbdpre), heat (configs[2]), lu (configs[4]).
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
RTOL = ATOL = 1e-5
TOUT = 1.0
HBM_NOMINAL = 8000.0     # GB/s, the figure north_star quotes
WORKLOADS = {
    # name: mx, my, jpre of the solve, jpre of the second (iteration-only) window, BASELINE configs index
    "web4096": dict(mx=4096, my=4096, jpre=3, jpre2=1, cfg=3),
    "web1024": dict(mx=1024, my=1024, jpre=1, jpre2=0, cfg=1),
}


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
def cpu_sample(steps, warmup, mesh, jpre, mx_full, my_full):
    """Time the oracle (CPU restatement) on a bounded sample: the same food-web ns=20 problem, same preconditioner,
    same tolerances, on a `mesh` x `mesh` sub-mesh, W+K steps, NLI/s over the last K steps scaled to the full mesh by
    the unknown ratio (every sweep of the path is O(N))."""
    from tests import oracle_py

    orc = oracle_py.load()
    total = steps + warmup
    import ctypes as C

    logbuf = np.zeros((total + 8, 3))
    orc.L.o_set_steplog.argtypes = [C.c_void_p, C.c_long]
    orc.L.o_set_steplog(logbuf.ctypes.data_as(C.c_void_p), total + 8)
    t0 = time.time()
    r = orc.run_web(NP_SPECIES, mesh, mesh, [TOUT], jpre=jpre, rtol=RTOL, atol=ATOL, max_steps=total)
    wall = time.time() - t0
    n = int(orc.L.o_steplog_count())
    orc.L.o_set_steplog(None, 0)
    if n < total or r["idid"] < 0:
        raise RuntimeError("oracle sample run ended early: idid=%d steps=%d" % (r["idid"], n))
    w = max(warmup, 1)
    dt = logbuf[total - 1, 0] - logbuf[w - 1, 0]
    dnli = logbuf[total - 1, 1] - logbuf[w - 1, 1]
    frac = (mesh * mesh) / float(mx_full * my_full)
    rate_sample = dnli / dt
    return {
        "value": rate_sample * frac, "unit": "it/s", "cores": 1, "kind": "port",
        "sample": "oracle (C++ restatement of the Fortran reference; no Fortran compiler in the image), food web ns=20, "
                  "jpre=%d, on a %dx%d sub-mesh, steps %d..%d of the same run: %.0f GMRES its in %.2f s = %.3f it/s at "
                  "sample size; value = that x %.6g (unknown ratio to the %dx%d mesh: every sweep of the path is O(N))"
                  % (jpre, mesh, mesh, w + 1, total, dnli, dt, rate_sample, frac, mx_full, my_full),
        "scale_factor": frac, "rate_at_sample_size": rate_sample,
        "nproc": os.cpu_count(), "wall_s": wall, "nli": float(dnli), "ms_per_step": 1e3 * dt / max(total - w, 1),
    }


def cpu_sample_native(steps, warmup, mesh, jpre, mx_full, my_full):
    """Courtesy number: the same oracle sources rebuilt on THIS host with -O3 -march=native
    (FMA contraction allowed, so not the parity arithmetic) and timed on the same sample in a child process."""
    import tempfile

    odir = os.path.join(ROOT, "oracle")
    srcs = [os.path.join(odir, f) for f in ("blas_linpack.cpp", "krylov.cpp", "rbdpre_problems.cpp", "daskr_driver.cpp")]
    with tempfile.TemporaryDirectory() as tmp:
        lib = os.path.join(tmp, "liboracle.so")
        subprocess.run(["g++", "-O3", "-march=native", "-fPIC", "-std=c++17", "-shared", "-o", lib] + srcs + ["-lm"], check=True)
        code = ("import sys, json; sys.path.insert(0, %r); import bench; from tests import oracle_py; oracle_py.LIB = %r; "
                "oracle_py.ORACLE_DIR = %r; r = bench.cpu_sample(%d, %d, %d, %d, %d, %d); print(json.dumps({k: r[k] for k in ('value', 'nli', 'wall_s')}))"
                % (ROOT, lib, tmp, steps, warmup, mesh, jpre, mx_full, my_full))
        out = subprocess.run([sys.executable, "-c", code], check=True, stdout=subprocess.PIPE, text=True).stdout
    return json.loads(out.strip().splitlines()[-1])


def workload_text(wl, mx, my, jpre, myloc=None):
    pre = {1: "drbdpre reaction block-diagonal preconditioner (jpre=1)",
           3: "preconditioner as shipped by example_web (jpre=3: Gauss-Seidel spatial factor x drbdpre reaction block factor)"}.get(jpre, "jpre=%d" % jpre)
    s = "food-web DAE ns=20 (np=10), %dx%d mesh of the unit square" % (mx, my)
    if myloc is not None and myloc != my:
        s += " split into y-slabs of %d rows per GPU" % myloc
    return s + ", %s, rtol=atol=1e-5, IC calculation on, tout=1 (BASELINE configs[%d])" % (pre, WORKLOADS.get(wl, {}).get("cfg", -1))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    w = WORKLOADS[args.workload]
    mx, my = args.mx or w["mx"], args.my or w["my"]
    jpre = args.jpre or w["jpre"]
    mesh = args.cpu_mesh or min(512, mx)
    cb = cpu_sample(args.steps, args.warmup, mesh, jpre, mx, my)
    if args.cpu_native:
        try:
            cb["sample"] += "; courtesy rebuild with -O3 -march=native on this host: %.4f it/s" % cpu_sample_native(args.steps, args.warmup, mesh, jpre, mx, my)["value"]
        except Exception as ex:
            cb["sample"] += "; courtesy -O3 -march=native run failed: %r" % (ex,)
    line = {
        "impl": "reference", "metric": "gmres_iters_per_s", "value": cb["value"], "unit": "it/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"] / cb["scale_factor"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_text(args.workload, mx, my, jpre) + "; CPU arm = the serial reference algorithm (C++ restatement) on a "
                               "%dx%d sub-mesh sample, rate scaled by the unknown ratio %.6g" % (mesh, mesh, cb["scale_factor"]),
                   "jpre": jpre, "cpu_mesh": mesh, "scale_factor": cb["scale_factor"]},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "nproc", "scale_factor", "rate_at_sample_size")},
        "e2e": {"value": cb["value"], "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------- our arm
def b_it(n, ns, ll):
    """algorithmic bytes of one Arnoldi iteration at basis index ll (phase A: read v, wght, y, ydot, savr,
    LU block, pivots, write v_{ll+1}; phase B: modified Gram-Schmidt + normalisation)."""
    return n * (8 * (6 + ns) + 4) + n * 8 * (4 * ll + 3)


def run_web_bench(args, out_fd):
    import torch
    import torch.distributed as dist

    import daskr_b200 as dk

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
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dk.comm_init_from_torch(rank, world)
        warm = torch.zeros(1, device="cuda")
        dist.all_reduce(warm)
        torch.cuda.synchronize()

    w = WORKLOADS[args.workload]
    K, W = args.steps, args.warmup
    ns = 2 * NP_SPECIES
    mx, my = args.mx or w["mx"], args.my or w["my"]
    jpre = args.jpre or w["jpre"]
    jpre2 = w["jpre2"] if args.jpre2 is None else args.jpre2
    if my % world != 0:
        log("note: %d rows do not divide evenly over %d ranks; slabs differ by one row" % (my, world))
    probe = np.zeros(64, dtype=np.int32)
    dk.setup_rbdpre(mx, my, ns, NP_SPECIES, 0, probe)
    jy0, myloc = dk.slab()
    neq = ns * mx * myloc
    neq_g = ns * mx * my
    iwork = np.zeros(40 + neq, dtype=np.int32)
    dk.setup_rbdpre(mx, my, ns, NP_SPECIES, 40, iwork)
    dk.web_setpar(NP_SPECIES, mx, my)
    stream = torch.cuda.ExternalStream(dk.lib().dkb_stream())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxrank(x):
        tm = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        return float(tm.item())

    # ---- consistent initial values (untimed set-up; host arrays) ----
    info = np.zeros(20, dtype=np.int32)
    info[10] = 1; info[11] = 1; info[13] = 1; info[14] = 1; info[15] = 1
    rwork = np.zeros(60)
    rpar = np.zeros(1)
    ipar = np.array([jpre, 0], dtype=np.int32)
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
    ic_counters = {k: int(iwork[i]) for k, i in (("nre", 11), ("nje", 12), ("nni", 18), ("nli", 19), ("nps", 20))}
    log("[rank %d] IC calculation %.2fs, counters %s" % (rank, time.time() - t0, ic_counters))
    d_y = torch.from_numpy(c).cuda()
    d_yp = torch.from_numpy(cdot).cuda()
    d_y_ic, d_yp_ic = d_y.clone(), d_yp.clone()
    info[10] = 0

    def events():
        return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    CNT = (("nst", 10), ("nre", 11), ("nje", 12), ("netf", 13), ("ncfn", 14), ("ncfl", 15), ("nni", 18), ("nli", 19), ("nps", 20))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

    # =========== (1) device-resident: W warm-up steps, K timed steps, one call each ===========
    dk.set_option(dk.OPT_DEVICE_IO, 1)
    info[0] = 0
    info[2] = 0
    state = {"t": 0.0, "failed": None}

    def advance(nsteps, allow_fail=False):
        """one dkb_daskr call that takes exactly nsteps BDF steps (per-call step limit -> idid=-1)"""
        dk.set_option(dk.OPT_MAX_STEPS, nsteps)
        nst0 = int(iwork[10]) if info[0] != 0 else 0
        state["t"], idid = dk.daskr(dk.WEB_RES, neq, state["t"], d_y.data_ptr(), d_yp.data_ptr(), TOUT, info, RTOL, ATOL, rwork, iwork,
                                    rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)
        if idid == -1:
            info[0] = 1   # "too many steps on this call": continue (DASKR doc, idid=-1)
        elif idid < 0:
            if not allow_fail:
                raise SystemExit("daskr failed in the benchmark window: idid=%d at t=%g" % (idid, state["t"]))
            state["failed"] = idid
            info[0] = 1
        return int(iwork[10]) - nst0

    def timed_window(nwarm, nsteps, allow_fail=False):
        if nwarm > 0:
            advance(nwarm, allow_fail)
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
        taken = advance(nsteps, allow_fail)
        e1.record(stream)
        barrier()
        w1 = time.time()
        clocks = sampler.stop(w0, w1)
        ms = maxrank(e0.elapsed_time(e1))
        prof = dk.prof_get()
        hist = dk.krylov_stats(16)
        launches = dk.launch_count()
        dk.prof_enable(False)
        cnt1 = iwork[:40].copy()
        dcount = {k: int(cnt1[i] - cnt0[i]) for k, i in CNT}
        return dict(ms=ms, taken=taken, prof=prof, hist=hist, launches=launches, clocks=clocks, dcount=dcount,
                    counters_end={k: int(cnt1[i]) for k, i in CNT})

    def iteration_roofline(win, with_gs):
        """whole Arnoldi iteration: B_it(ll) over the ll histogram / event-timed iterations, per GPU"""
        kit = win["prof"].get("krylov_iter", {"ms": 0.0, "count": 0})
        hist = win["hist"]
        nit = int(hist.sum())
        if nit == 0 or kit["ms"] <= 0:
            return None
        byts = float(sum(int(hist[l - 1]) * b_it(neq, ns, l) for l in range(1, len(hist) + 1)))
        ach = byts / (kit["ms"] * 1e-3) / 1e9
        r = {"bytes_formula": "SURVEY 8(d): B_it(ll) = N[8(6+ns)+4] + 8N(4ll+3), N = unknowns of this GPU's slab",
             "ll_hist": [int(x) for x in hist[:8]], "iterations": nit, "mean_ll": float(sum((l + 1) * int(hist[l]) for l in range(len(hist))) / nit),
             "ms_per_iteration": kit["ms"] / nit, "achieved": ach, "unit": "GB/s per GPU",
             "frac_of_measured_peak": ach / peak, "frac_of_nominal_8TBs": ach / HBM_NOMINAL,
             "share_of_window": kit["ms"] / win["ms"]}
        if with_gs:
            r["note"] = ("B_it has no term for the Gauss-Seidel factor of jpre=3 (it works on data the iteration already holds); "
                         "counting its minimum of one read + one write of the vector (16 B per unknown) gives frac_with_gs_16B")
            ach2 = (byts + 16.0 * neq * nit) / (kit["ms"] * 1e-3) / 1e9
            r["frac_with_gs_16B_of_measured"] = ach2 / peak
        return r

    def kernel_roofline(win, jp):
        """the kernel with the largest event-timed total in the window"""
        P = win["prof"]
        cand = []
        one = P.get("datv_single_kernel", {"ms": 0.0, "count": 0})
        if one["count"]:
            cand.append(("web_datv_single_kernel<20> (fused datv: perturbed-state residual + psol_rbdpre block solve)",
                         one, neq * (8 * (6 + ns) + 4), "web_datv_single_ns20"))
        tpp = P.get("tpp_psol_kernel", {"ms": 0.0, "count": 0})
        if tpp["count"]:
            cand.append(("tpp_psol_kernel<20> (psol_rbdpre block solve + datv's '* wght')", tpp, neq * (8 * ns + 4 + 8 * 3), "tpp_psol_ns20"))
        gs = P.get("gauss_seidel", {"ms": 0.0, "count": 0})
        if gs["count"]:
            cand.append(("Gauss-Seidel factor (5 sweeps of example_web.f90:393-575 in one streaming kernel; algorithmic = z read + z written)",
                         gs, neq * 16, "gs_ns20"))
        rd = P.get("resdatv_kernel", {"ms": 0.0, "count": 0})
        if rd["count"]:
            cand.append(("web_datv_single_kernel<20,MODE 1> (perturbed-state residual - savr)", rd, neq * 56, "resdatv_ns20"))
        if not cand:
            return {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None}
        cand.sort(key=lambda x: -x[1]["ms"])
        out = None
        others = {}
        for name, slot, byts, key in cand:
            avg = slot["ms"] / slot["count"]
            ach = byts / (avg * 1e-3) / 1e9
            ent = {"kernel": name, "launches": slot["count"], "avg_launch_ms": avg, "algorithmic_bytes_per_launch": byts,
                   "achieved": ach, "frac": ach / peak, "share_of_window": slot["ms"] / win["ms"]}
            if out is None:
                out = {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                       "peak_source": peak_src, "algorithmic_bytes_per_launch": byts, "launches": slot["count"], "avg_launch_ms": avg,
                       "share_of_step_time": slot["ms"] / win["ms"]}
                try:
                    tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
                    k2 = "%s_%dx%d" % (key, mx, myloc)
                    if k2 in tr:
                        out["traffic"] = tr[k2]["dram_bytes_per_launch"]
                        out["traffic_source"] = tr[k2].get("source")
                except Exception:
                    pass
            else:
                others[key] = ent
        out["other_kernels"] = others
        # K5 (jac_rbdpre: FD block + LU) is the one compute-shaped kernel of the path: report its FP64 rate.
        jc = P.get("jac", {"ms": 0.0, "count": 0})
        if jc["count"] > 0 and jc["ms"] > 0:
            fl_pt = (ns + 1) * (2 * ns * ns + 3 * ns) + sum(m + 2 * m * m for m in range(1, ns))
            out["jac_kernel"] = {"kernel": "web_jac3_kernel<20> (FD block Jacobian + LU, two rows of a block per lane; incl. the first-failure flag readback)",
                                 "calls": jc["count"], "avg_ms": jc["ms"] / jc["count"], "flops_per_point": fl_pt,
                                 "achieved_gflops": fl_pt * (neq // ns) / (jc["ms"] / jc["count"] * 1e-3) / 1e9,
                                 "fp64_peak_gflops": {"dfma": 33600.0, "dmul_dadd_pairs_no_fma": 18400.0,
                                                      "source": "tools/fp64_peak.cu on this GPU type (profiles/README.md)"},
                                 "share_of_step_time": jc["ms"] / win["ms"]}
        return out

    win = timed_window(W, K)
    if win["taken"] != K:
        raise SystemExit("timed window took %d steps instead of %d" % (win["taken"], K))
    ms = win["ms"]
    dnli = win["dcount"]["nli"]
    value = dnli / (ms * 1e-3)
    t_end_window = state["t"]
    log("[rank %d] device-resident: %d steps in %.1f ms, counters %s, t=%.3e h=%.2e" % (rank, K, ms, win["dcount"], state["t"], rwork[6]))
    roof = kernel_roofline(win, jpre)
    roof["whole_iteration"] = iteration_roofline(win, jpre == 3)
    phase_ms = {k: round(v["ms"], 3) for k, v in win["prof"].items()}
    sync_gap_ms = ms - sum(win["prof"].get(k, {"ms": 0.0})["ms"] for k in ("krylov_iter", "psol0", "jac", "res", "vec", "norm"))

    # digest of the state at the end of the timed window: global mean(c1), wrms(y) with DASKR's weights, counters
    yv = d_y.view(-1, ns)
    sums = torch.stack([yv[:, 0].sum(), ((d_y / (RTOL * d_y.abs() + ATOL)) ** 2).sum()])
    if world > 1:
        dist.all_reduce(sums)
    digest = {"mean_c1": float(sums[0].item()) / (mx * my), "wrms_y": float((sums[1].item() / neq_g) ** 0.5), "t": state["t"], "h": float(rwork[2]),
              "counters": win["counters_end"], "ic_counters": ic_counters}

    # =========== (1b) the drbdpre iteration on the same run: ipar(1) = 1 ===========
    drb = None
    if jpre2 and jpre2 != jpre:
        ipar[0] = jpre2
        try:
            win2 = timed_window(args.warmup2, args.steps2, allow_fail=True)
            it2 = iteration_roofline(win2, jpre2 == 3)
            kr2 = kernel_roofline(win2, jpre2)
            drb = {"jpre": jpre2, "what": "same run continued with ipar(1)=%d (drbdpre reaction block factor alone): %d warm-up + %d timed steps"
                                          % (jpre2, args.warmup2, args.steps2),
                   "steps_taken": win2["taken"], "ms": win2["ms"], "counters": win2["dcount"], "nli_per_s": win2["dcount"]["nli"] / (win2["ms"] * 1e-3),
                   "idid_failure": state["failed"], "whole_iteration": it2,
                   "kernel": {k: kr2.get(k) for k in ("kernel", "avg_launch_ms", "achieved", "frac", "algorithmic_bytes_per_launch", "launches", "traffic")},
                   "phase_ms": {k: round(v["ms"], 3) for k, v in win2["prof"].items()}}
            log("[rank %d] drbdpre window: %s" % (rank, json.dumps(drb)[:600]))
        except SystemExit as ex:
            drb = {"jpre": jpre2, "error": str(ex)}
        ipar[0] = jpre

    # =========== (2) e2e: the first window through the host-buffer API, one call per step ===========
    dk.set_option(dk.OPT_DEVICE_IO, 0)
    info[0] = 0
    info[2] = 1   # info(3)=1: return after every step (y, yprime copied to the host each time)
    tcur = 0.0
    c[:] = d_y_ic.cpu().numpy()
    cdot[:] = d_yp_ic.cpu().numpy()
    del d_y_ic, d_yp_ic
    dk.set_option(dk.OPT_MAX_STEPS, 500)
    dk.set_option(dk.OPT_ASYNC_OUT, 0 if args.sync_out else 1)
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
    dk.output_wait()   # the last step's y / yprime have arrived in the pinned host arrays
    e1.record(stream)
    barrier()
    dk.set_option(dk.OPT_ASYNC_OUT, 0)
    ms_e2e = maxrank(e0.elapsed_time(e1))
    dnli_e2e = int(iwork[19]) - nli0
    hdr_bytes = 60 * 8 + 40 * 4 + 20 * 4 + 5 * 8
    e2e = {"value": dnli_e2e / (ms_e2e * 1e-3), "unit": "it/s", "h2d_bytes_per_step": hdr_bytes, "d2h_bytes_per_step": 16 * neq,
           "h2d_bytes_initial": 16 * neq, "initial_call_ms": initial_call_ms, "ms_per_step": ms_e2e / K, "nli": dnli_e2e,
           "bytes_are": "per GPU (each rank copies its slab)",
           "mode": "pinned host y/yprime, info(3)=1: one dkb_daskr call per step, y+yprime D2H at every return%s; "
                   "y/yprime are inputs only on the initial call (DASKR convention), so per-step H2D is the scalar headers"
                   % ("" if args.sync_out else " (DKB_OPT_ASYNC_OUT: the transfer of step n runs on a second stream under step n+1; "
                                               "the window ends when the last transfer has arrived)")}
    log("[rank %d] e2e: %d steps in %.1f ms, dNLI=%d" % (rank, K, ms_e2e, dnli_e2e))

    # =========== (3) CPU baseline (rank 0, N=1 only) ===========
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cb = cpu_sample(min(K, 20), min(W, 20), args.cpu_mesh or 256, jpre, mx, my)
            cpu = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "nproc", "scale_factor", "rate_at_sample_size")}
        except Exception as ex:   # the CPU leg must never take the GPU number down with it
            cpu = {"value": None, "unit": "it/s", "cores": 1, "kind": "port", "sample": "failed: %r" % (ex,)}

    if rank == 0:
        line = {
            "metric": "gmres_iters_per_s", "value": value, "unit": "it/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_text(args.workload, mx, my, jpre, myloc),
                "neq_global": neq_g, "neq_per_gpu": neq, "domain": "[0,1]x[0,1]", "jpre": jpre, "gmres": "MAXL=5 KMP=5 NRMAX=5 EPLI=0.05 (DASKR defaults)",
                "l2": "inputs larger than L2: %.0f MB per length-N vector per GPU, %.2f GB of LU blocks streamed per Krylov iteration per GPU" % (8 * neq / 1e6, 8.0 * ns * neq / 1e9),
                "window": "steps %d..%d after the IC calculation, t_end=%.3e" % (W + 1, W + K, t_end_window),
                "counters": win["dcount"],
            },
            "e2e": e2e, "gpu_launches": win["launches"], "clocks": win["clocks"], "roofline": roof, "phase_ms": phase_ms,
            "host_gap_ms": round(sync_gap_ms, 3), "digest": digest,
        }
        if drb is not None:
            line["drbdpre_iteration"] = drb
        if cpu is not None:
            line["cpu_baseline"] = cpu
        os.write(out_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="web4096", choices=sorted(WORKLOADS) + ["heat", "lu"])
    ap.add_argument("--mx", type=int, default=None)
    ap.add_argument("--my", type=int, default=None, help="GLOBAL mesh rows (split over the GPUs)")
    ap.add_argument("--jpre", type=int, default=None, help="food-web preconditioner of the solve: 1 = drbdpre alone; 3 = as shipped")
    ap.add_argument("--jpre2", type=int, default=None, help="preconditioner of the second window (0: skip)")
    ap.add_argument("--steps2", type=int, default=6)
    ap.add_argument("--warmup2", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--sync-out", action="store_true", help="e2e leg: blocking y/yprime transfer at every return (no DKB_OPT_ASYNC_OUT)")
    ap.add_argument("--cpu-native", action="store_true",
                    help="--impl reference only: also time a -O3 -march=native rebuild of the oracle (courtesy number)")
    ap.add_argument("--cpu-mesh", type=int, default=None)
    ap.add_argument("--heat-atol", type=float, default=1e-7, help="--workload heat: absolute tolerance (see tools/bench_extra.py)")
    ap.add_argument("--lu-full", action="store_true", help="--workload lu: include the 1e7-block entries (up to 82 GB of blocks)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("note: W >= 3 warm-up steps are required for a valid number; got %d" % args.warmup)
    if args.impl == "reference":
        if args.workload in ("heat", "lu"):
            from tools import bench_extra
            return bench_extra.run_reference(args)
        return run_reference(args)
    # Everything but the one JSON line goes to stderr: NCCL (NCCL_DEBUG=INFO is left alone so that the driver can check the
    # communicator) and libraries print to fd 1 at will.
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)
    if args.workload in ("heat", "lu"):
        from tools import bench_extra
        return bench_extra.run(args, out_fd)
    return run_web_bench(args, out_fd)


if __name__ == "__main__":
    sys.exit(main())
