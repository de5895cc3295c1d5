"""daskr_b200 -- B200-native Krylov path of DASKR behind DASKR's own calling convention.

Host-side mirror of the reference interface (HugoMVale/daskr):

    setup_rbdpre(mx, my, ns, nsd, lid, iwork)          src/daskr_rbdpre.f90:102
    daskr(res, neq, t, y, ydot, tout, info, rtol, atol,
          rwork, iwork, rpar, ipar, jac, psol)          src/daskr.f:1

with the same argument meaning, `info/rwork/iwork` conventions, `idid` codes
and counters.  All arithmetic runs in hand-written sm_100a CUDA kernels inside
`_lib/libdaskr_b200.so` (C ABI: include/daskr_b200.h); there is no CPU
fallback -- importing works without a GPU, computing does not.
"""
from .api import (  # noqa: F401
    DaskrError,
    lib,
    init,
    finalize,
    set_option,
    setup_rbdpre,
    setup_rbgpre,
    slab,
    web_setpar,
    web_cinit,
    heat_setpar,
    heat_uinit,
    daskr,
    WEB_RES,
    WEB_JAC,
    WEB_PSOL,
    HEAT_RES,
    HEAT_JAC,
    HEAT_PSOL,
    WEB_RT,
    HEAT_RT,
    state_export,
    state_import,
    dgefa_batched,
    dgesl_batched,
    solve_web,
    solve_heat,
    prof_enable,
    prof_reset,
    prof_get,
    launch_count,
    krylov_stats,
    comm_init_from_torch,
    OPT_DEVICE_IO,
    OPT_FUSED,
    OPT_MAX_STEPS,
    OPT_ASYNC_OUT,
    output_wait,
)

__version__ = "0.1.0"
