import numpy as np, sys
sys.path.insert(0,'/root/repo')
import daskr_b200 as dk
from tests import oracle_py
o=oracle_py.load(); dk.init(0)
for ns in (3,5,6):
    rng=np.random.default_rng(0)
    for nblk in (2,64,65,1237):
        bd=rng.standard_normal((nblk,ns*ns))
        lu_g,ip_g,info_g,first=dk.dgefa_batched(bd,ns)
        lu_o,ip_o,info_o=o.dgefa_batched(bd,ns)
        badp=np.nonzero((ip_g!=ip_o).any(axis=1))[0]
        print("ns",ns,"nblk",nblk,"pivot mismatches",badp.size, badp[:10], "lu equal", np.array_equal(lu_g,lu_o))
        if badp.size:
            p=badp[0]
            print(" in ",bd[p]); print(" gpu",lu_g[p],ip_g[p]); print(" cpu",lu_o[p],ip_o[p])
