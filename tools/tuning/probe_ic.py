"""Which meshes does the consistent-initialisation Newton (info(11)=1) of the ns=20 food web survive with the reaction
block factor alone (jpre=1)?  Prints idid and counters of the IC call per mesh (GPU)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import daskr_b200 as dk

dk.init(0)
for m in [int(a) for a in sys.argv[1:]]:
    for jpre in (1, 3):
        t0 = time.time()
        r = dk.solve_web(10, m, m, [1e-8], jpre=jpre, max_steps=1, y_out=False)
        print("mesh %d jpre %d: idid %d counters(nst,nre,nje,nni,nli,ncfn,ncfl)=%s %.1fs"
              % (m, jpre, r["idid"], r["counters"][-1][[10, 11, 12, 18, 19, 14, 15]].tolist(), time.time() - t0), flush=True)
