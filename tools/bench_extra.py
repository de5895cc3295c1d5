#!/usr/bin/env python
"""Extra measurements for the BASELINE configs that are not the driver's bench line:
   config 3  heat equation 8192x8192, rbdpre ns=1, Krylov path
   config 5  batched LU micro-sweep ns in {2,4,8,16,20,32} x 1e4..1e7 blocks
Writes one JSON document to stdout (kept under profiles/).  Timings are host wall clock around
C-ABI calls that end with a stream synchronise (one kernel launch each), inputs resident in HBM.
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import daskr_b200 as dk  # noqa: E402
from daskr_b200.api import DeviceBuffer  # noqa: E402

dk.init(0)
L = dk.lib()
peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
out = {"gpu": "B200", "hbm_peak_gbs": json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else None}

# ---------------- config 5: LU micro-sweep on the reference bd/ipbd layout
sweep = []
rng = np.random.default_rng(0)
for ns in (2, 4, 8, 16, 20, 32):
    for nblk in (10_000, 100_000, 1_000_000, 10_000_000):
        # up to 81.9 GB of blocks (ns=32 x 1e7): a base batch of <= 1e6 random blocks is generated on the host and
        # tiled over the device array (every rep restores the array from the base: dgefa works in place)
        base_n = min(nblk, 1_000_000)
        ntile = nblk // base_n
        bd = rng.standard_normal((base_n, ns * ns))
        rhs = rng.standard_normal((base_n, ns))
        d_bd0, d_b0 = DeviceBuffer.from_host(bd), DeviceBuffer.from_host(rhs)
        bd_bytes, rhs_bytes = bd.nbytes, rhs.nbytes
        d_bd, d_b, d_ip = DeviceBuffer(bd_bytes * ntile), DeviceBuffer(rhs_bytes * ntile), DeviceBuffer(4 * ns * nblk)
        del bd, rhs

        def restore():
            for k in range(ntile):
                L.dkb_memcpy_d2d(C.c_void_p(d_bd.ptr.value + k * bd_bytes), d_bd0.ptr, bd_bytes)
                L.dkb_memcpy_d2d(C.c_void_p(d_b.ptr.value + k * rhs_bytes), d_b0.ptr, rhs_bytes)
            L.dkb_sync()

        first = C.c_int64(0)
        reps = 3 if nblk >= 10_000_000 else 5 if nblk >= 1_000_000 else 20
        tf = ts = 0.0
        for r in range(reps + 1):
            restore()
            t0 = time.perf_counter()
            L.dkb_dgefa_batched(d_bd.ptr, d_ip.ptr, ns, nblk, None, C.byref(first))
            t1 = time.perf_counter()
            L.dkb_dgesl_batched(d_bd.ptr, d_ip.ptr, ns, nblk, d_b.ptr)
            t2 = time.perf_counter()
            if r > 0:
                tf += t1 - t0
                ts += t2 - t1
        tf /= reps
        ts /= reps
        fl_f = sum(m + 2 * m * m for m in range(1, ns))            # dgefa flops per block
        fl_s = 2 * ns * ns - ns
        # the last tile saw the same input as the first: same factors, pivots and solutions (64-bit indexing check)
        agree = True
        if ntile > 1:
            nchk = 1000
            for buf, per, dt in ((d_bd, ns * ns * 8, np.float64), (d_b, ns * 8, np.float64)):
                a0 = np.empty(nchk * per // 8, dtype=dt)
                a1 = np.empty_like(a0)
                L.dkb_memcpy_d2h(a0.ctypes.data_as(C.c_void_p), buf.ptr, a0.nbytes)
                off = (ntile - 1) * base_n * per
                L.dkb_memcpy_d2h(a1.ctypes.data_as(C.c_void_p), C.c_void_p(buf.ptr.value + off), a1.nbytes)
                agree = agree and bool(np.array_equal(a0, a1)) and bool(np.isfinite(a0).all())
        sweep.append({"ns": ns, "nblk": nblk, "gb_of_blocks": ns * ns * nblk * 8 / 1e9, "last_tile_equals_first": agree, "dgefa_ms": tf * 1e3, "dgesl_ms": ts * 1e3,
                      "dgefa_gflops": fl_f * nblk / tf / 1e9, "dgefa_gbs": (16 * ns * ns + 4 * ns) * nblk / tf / 1e9,
                      "dgesl_gflops": fl_s * nblk / ts / 1e9, "dgesl_gbs": (8 * ns * ns + 20 * ns) * nblk / ts / 1e9})
        print("lu ns=%d nblk=%d dgefa %.3f ms dgesl %.3f ms" % (ns, nblk, tf * 1e3, ts * 1e3), file=sys.stderr, flush=True)
        for b in (d_bd0, d_bd, d_ip, d_b0, d_b):
            b.free()
out["lu_sweep"] = sweep

# ---------------- config 3: heat 8192^2, first 40 steps (info(3)=1 stepping: y/yprime return to the host every step)
m = 8190
t0 = time.perf_counter()
r = dk.solve_heat(m, [0.01], max_steps=40, y_out=False)
dt = time.perf_counter() - t0
cnt = r["counters"][-1]
out["heat_8192"] = {"steps": int(cnt[10]), "nni": int(cnt[18]), "nli": int(cnt[19]), "nje": int(cnt[12]), "wall_s": dt,
                    "gmres_it_per_s_with_host_output_every_step": float(cnt[19]) / dt, "t_reached": r["t"],
                    "idid": r["idid"], "neq": (m + 2) ** 2}
print(json.dumps(out))
