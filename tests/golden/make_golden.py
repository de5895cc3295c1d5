"""Regenerates tests/golden/oracle_golden.json from the oracle (python tests/golden/make_golden.py).

These fixtures pin the ORACLE against accidental change.  They are not reference outputs: the
reference (Fortran) cannot be built in this image and ships no expected-output files for the
Krylov path -- parity with the reference itself stays unpinned.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import oracle_py  # noqa: E402

o = oracle_py.load()
out = {}
touts = [1e-8, 1e-6, 1e-4, 1e-2, 1.0]
r = o.run_web(1, 20, 20, touts)
out["web20"] = {
    "touts": touts,
    "c1_average": [float(y.reshape(20, 20, 2)[:, :, 0].mean()) for y in r["y"]],
    "counters_last": r["counters"][-1][[10, 11, 12, 18, 19, 20]].tolist(),
}
ht = [0.01 * 2 ** n for n in range(6)]
h = o.run_heat(10, ht)
out["heat10"] = {"touts": ht, "umax": np.abs(h["y"]).max(axis=1).tolist(), "counters_last": h["counters"][-1][[10, 18, 19]].tolist()}
rng = np.random.default_rng(0)
bd = rng.standard_normal((4, 25))
_, ip, _ = o.dgefa_batched(bd, 5)
out["dgefa5"] = {"ipvt": ip.tolist()}
# the shipped preconditioner settings: spatial Gauss-Seidel factor (jpre = 3), block grouping (5 x 5 groups), root
r3 = o.run_web(1, 20, 20, touts, jpre=3)
out["web20_jpre3"] = {"c1_average": [float(y.reshape(20, 20, 2)[:, :, 0].mean()) for y in r3["y"]],
                      "counters_last": r3["counters"][-1][[10, 11, 12, 18, 19, 20]].tolist()}
rg = o.run_web(1, 20, 20, touts, jpre=3, nxg=5, nyg=5)
out["web20_jpre3_groups5"] = {"c1_average": [float(y.reshape(20, 20, 2)[:, :, 0].mean()) for y in rg["y"]],
                              "counters_last": rg["counters"][-1][[10, 11, 12, 18, 19, 20]].tolist()}
from tests.test_gpu_roots import oracle_web_roots  # noqa: E402  (CPU helper living next to the GPU test)
rr = oracle_web_roots(o, 1, 20, 20, [1e-8, 1e-6, 1e-4, 1e-2, 1e-1])
out["web20_root"] = {"t": [t for t, _, _ in rr["roots"]], "jroot": [j.tolist() for _, j, _ in rr["roots"]],
                     "nst_at_root": [int(c[10]) for _, _, c in rr["roots"]]}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.json"), "w"), indent=1)
print("wrote oracle_golden.json")
