// driver.cu -- host side of DASKR's Krylov option above the device kernels.
//
// The north star keeps the BDF / Newton control logic (DASKR, ddstp, dnedk, dnsk
// and the initial-condition path ddasic/ddasik/dnsik/dlinsk/dfnrmk) as host code
// calling a thin layer over CUDA.  No Fortran compiler exists in this image, so
// that host logic is written here in C++ with the reference's scalar control
// flow, error codes and counters (src/daskr.f:1402-2493, 2911-3131;
// src/daskr_new.f90:149-567, 2004-3034) and with every whole-vector statement
//  replaced by a kernel from vecops.cu.  Scalars (norms, flags)
// come back to the host for the control decisions; vectors never do, except the
// y / yprime copies the DASKR calling convention itself requires at entry/exit.
//
// rwork keeps the reference's scalar header rwork(1:60); iwork keeps iwork(1:40)
// followed by the ICNSTR / ID arrays the caller provides.  The vector segments of
// rwork (DELTA, SAVR, E, WT, VT, PHI, WM incl. WP) live in device memory.
#include "dkb_internal.h"
#include <math.h>
#include <string.h>
#include <limits.h>
#include <algorithm>

// iwork pointers, src/daskr.f:1414-1420
enum { LML = 1, LMU = 2, LMTYPE = 4, LIWM = 1, LMXORD = 3, LJCALC = 5, LPHASE = 6, LK = 7, LKOLD = 8, LNS = 9,
       LNSTL = 10, LNST = 11, LNRE = 12, LNJE = 13, LETF = 14, LNCFN = 15, LNCFL = 16, LNIW = 17, LNRW = 18,
       LNNI = 19, LNLI = 20, LNPS = 21, LNPD = 22, LMITER = 23, LMAXL = 24, LKMP = 25, LNRMAX = 26, LLNWP = 27,
       LLNIWP = 28, LLOCWP = 29, LLCIWP = 30, LKPRIN = 31, LMXNIT = 32, LMXNJ = 33, LMXNH = 34, LLSOFF = 35,
       LNRTE = 36, LIRFND = 37, LICNS = 41 };
// rwork pointers, src/daskr.f:1424-1427
enum { LTSTOP = 1, LHMAX = 2, LH = 3, LTN = 4, LCJ = 5, LCJOLD = 6, LHOLD = 7, LS = 8, LROUND = 9, LEPLI = 10,
       LSQRN = 11, LRSQRN = 12, LEPCON = 13, LSTOL = 14, LEPIN = 15, LALPHA = 21, LBETA = 27, LGAMMA = 33,
       LPSI = 39, LSIGMA = 45, LT0 = 51, LTLAST = 52, LDELTA = 61,
       LR0 = 61 };   // R0, R1, RX (3*nrt words) follow the scalar header: the vector segments live on the device

static inline double dsign(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }

struct Tol {   // rtol / atol as DASKR receives them
    int iwt;
    double *rtol, *atol;
};

// ---------------------------------------------------------------- device allocation
static int alloc_solver(DkbCtx *c, i64 neq, int maxl, int kmp, int need_vt, int need_id, int need_icnstr, int iwt,
                        i64 user_lenwp, i64 user_leniwp)
{
    if (c->neq != neq) {   // per-unknown side arrays are sized by neq
        if (c->d_id) cudaFree(c->d_id);
        if (c->d_icnstr) cudaFree(c->d_icnstr);
        if (c->d_rtol) cudaFree(c->d_rtol);
        if (c->d_atol) cudaFree(c->d_atol);
        c->d_id = c->d_icnstr = nullptr;
        c->d_rtol = c->d_atol = nullptr;
    }
    {
        // every vector carries one mesh row of halo on each side (slab neighbours / stencil reach)
        const i64 pad = c->halo_pad;
        const i64 stride = ((neq + 31) / 32) * 32 + 2 * pad;
        int nvec = 2 /*y,yp*/ + 5 /*delta,savr,e,wt,vt*/ + 6 /*phi*/ + (maxl + 1) /*v*/ + 3 /*dl,z,wght*/;
        if (kmp < maxl) nvec += 1;   // wk separate from dl (dslvk:3118)
        if (!c->arena || c->arena_doubles != stride * nvec) {
            if (c->arena) cudaFree(c->arena);
            c->arena = nullptr;
            c->arena_doubles = stride * nvec;
            CUDA_TRY(cudaMalloc(&c->arena, sizeof(double) * c->arena_doubles));
        }
        CUDA_TRY(cudaMemsetAsync(c->arena, 0, sizeof(double) * c->arena_doubles, c->stream));
        double *p = c->arena + pad;
        auto take = [&]() { double *r = p; p += stride; return r; };
        c->y = take();
        c->yp = take();
        c->delta = take();
        c->savr = take();
        c->e = take();
        c->wt = take();
        c->vt = take();
        for (int j = 0; j < 6; ++j) c->phi[j] = take();
        for (int j = 0; j <= maxl; ++j) c->v[j] = take();
        c->dl = take();
        c->z = take();
        c->wght = take();
        c->wk = (kmp < maxl) ? take() : c->dl;
        // IC scratch overlays the start of WM (daskr.f:1842-1844)
        c->yic = c->v[0];
        c->ypic = c->v[1];
        // three distinct vectors (daskr.f:1842-1844); with maxl = 1 the basis has only two, so pwk takes the datv scratch
        c->pwk = (maxl >= 2) ? c->v[2] : c->z;
        c->neq = neq;
        c->maxl = maxl;
        c->kmp = kmp;
        // several ranks on one node: map the neighbours' arenas (same layout everywhere) for the in-kernel halo exchange
        if (c->nranks > 1) peer_map_arena();
    }
    (void)need_vt;
    if (need_id && !c->d_id) CUDA_TRY(cudaMalloc(&c->d_id, sizeof(int) * neq));
    if (need_icnstr && !c->d_icnstr) CUDA_TRY(cudaMalloc(&c->d_icnstr, sizeof(int) * neq));
    if (iwt && !c->d_rtol) {
        CUDA_TRY(cudaMalloc(&c->d_rtol, sizeof(double) * neq));
        CUDA_TRY(cudaMalloc(&c->d_atol, sizeof(double) * neq));
    }
    // preconditioner storage: rbdpre blocks/pivots (64-bit sizes) plus whatever the caller asked
    // for through iwork(27:28)
    {
        i64 lenwp = user_lenwp, leniwp = user_leniwp;
        if (c->rbg_set) {   // block grouping: one block per group, reference layout
            lenwp += (i64)c->ngx * c->ngy * c->ns * c->ns;
            leniwp += (i64)c->ngx * c->ngy * c->ns;
        } else if (c->rbd_set) {   // interleaved layout pads the point count to a multiple of 32
            lenwp += tpp_lut_doubles(c->ns, c->npts);
            leniwp += tpp_pt_ints(c->ns, c->npts);
        }
        if (lenwp > c->lenwp) {
            if (c->wp) cudaFree(c->wp);
            CUDA_TRY(cudaMalloc(&c->wp, sizeof(double) * std::max<i64>(lenwp, 1)));
            c->lenwp = lenwp;
        }
        if (leniwp > c->leniwp) {
            if (c->iwp) cudaFree(c->iwp);
            CUDA_TRY(cudaMalloc(&c->iwp, sizeof(int) * std::max<i64>(leniwp, 1)));
            c->leniwp = leniwp;
        }
    }
    return 0;
}

// The solver workspace without a dkb_daskr call: what a host that keeps DASKR's own driver (the Fortran ddstp / dnedk /
// dnsk of the north star) allocates once before it calls dkb_dslvk and the vector-op layer.  lenwp / leniwp: extra
// preconditioner storage in doubles / ints on top of what dkb_setup_rbdpre / rbgpre imply (iwork(27:28) of the reference).
extern "C" int dkb_alloc_workspace(int64_t neq, int maxl, int kmp, int64_t lenwp, int64_t leniwp)
{
    DkbCtx *c = g_ctx;
    if (!c) return dkb_fail(DKB_E_STATE, "dkb_init not called");
    if (neq <= 0 || maxl < 1 || maxl > DKB_MAXL_MAX || kmp < 1 || kmp > maxl)
        return dkb_fail(DKB_E_ARG, "dkb_alloc_workspace: need neq > 0, 1 <= kmp <= maxl <= %d", DKB_MAXL_MAX);
    if (c->rbd_set && neq != (i64)c->ns * c->npts)
        return dkb_fail(DKB_E_ARG, "dkb_alloc_workspace: neq does not match the setup_rbdpre slab (ns*mx*myloc)");
    int r = alloc_solver(c, neq, maxl, kmp, 1, 1, 0, 0, lenwp, leniwp);
    if (r != 0) return r;
    c->neq_g = neq;
    if (c->nranks > 1) {
        double ng = (double)neq;
        cudaMemcpyAsync(c->d_scal + 50, &ng, sizeof(double), cudaMemcpyHostToDevice, c->stream);
        allreduce_sum(50, 1);
        scal_fetch(50, 1, &ng);
        c->neq_g = (i64)(ng + 0.5);
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---------------------------------------------------------------- small helpers
static int ewt_set(DkbCtx *c, const Tol &tol, const double *y, bool with_vt)
{
    // ddawts + dinvwt (+ vt mask).  Returns the first non-positive weight (1-based) or 0.
    c->h_iflag[0] = INT_MAX;
    cudaMemcpyAsync(c->d_iflag, c->h_iflag, sizeof(int), cudaMemcpyHostToDevice, c->stream);
    v_ewt(c->neq, tol.iwt, tol.rtol[0], tol.atol[0], c->d_rtol, c->d_atol, y, c->wt, with_vt ? c->vt : nullptr,
          c->d_id, c->d_iflag);
    fetch_words(c->d_iflag, c->h_iflag, sizeof(int));
    dkb_cuda_check("ewt_set");
    int bad = c->h_iflag[0];
    if (c->nranks > 1) {
        // every rank must take the same branch
        double flag = (bad != INT_MAX) ? 1.0 : 0.0;
        cudaMemcpyAsync(c->d_scal + 50, &flag, sizeof(double), cudaMemcpyHostToDevice, c->stream);
        allreduce_max(50, 1);
        scal_fetch(50, 1, &flag);
        if (flag > 0.0 && bad == INT_MAX) bad = 1;
    }
    return bad == INT_MAX ? 0 : bad;
}

static void interp(DkbCtx *c, double t, double tout, int kold, const double *psi /*1-based base*/)
{
    // ddatrp, src/daskr_new.f90:782-826: coefficients on the host, one fused multi-vector kernel
    double cc[6] = {1, 0, 0, 0, 0, 0}, dd[6] = {0, 0, 0, 0, 0, 0};
    const double dt = tout - t;
    double cv = 1.0, dv = 0.0, gama = dt / psi[1];
    for (int j = 2; j <= kold + 1; ++j) {
        dv = dv * gama + cv / psi[j - 1];
        cv = cv * gama;
        gama = (dt + psi[j - 1]) / psi[j];
        cc[j - 1] = cv;
        dd[j - 1] = dv;
    }
    v_predict(c->neq, kold + 1, c->phi, cc, dd, c->y, c->yp);
}

// DKB_OPT_ASYNC_OUT: y / yprime leave through a device-side snapshot and a second stream, so that the PCIe transfer of
// this return overlaps the next call's time step (the caller's arrays are complete after dkb_output_wait(), or once
// the next dkb_daskr call has returned).  The snapshot costs 32 bytes per unknown of HBM traffic instead of ~16 bytes
// per unknown of PCIe time on the solver's stream.
static bool copy_out_async(DkbCtx *c, double *y, double *yprime)
{
    const size_t nb = sizeof(double) * c->neq;
    if (!c->out_stream) {
        if (cudaStreamCreateWithFlags(&c->out_stream, cudaStreamNonBlocking) != cudaSuccess) return false;
        cudaEventCreateWithFlags(&c->ev_snap, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->ev_out, cudaEventDisableTiming);
    }
    if (c->out_cap < c->neq) {
        cudaStreamSynchronize(c->out_stream);
        cudaFree(c->out_y);
        cudaFree(c->out_yp);
        c->out_y = c->out_yp = nullptr;
        c->out_cap = 0;
        if (cudaMalloc(&c->out_y, nb) != cudaSuccess || cudaMalloc(&c->out_yp, nb) != cudaSuccess) {
            cudaFree(c->out_y);
            c->out_y = nullptr;
            cudaGetLastError();
            return false;
        }
        c->out_cap = c->neq;
    }
    if (c->out_pending) cudaStreamWaitEvent(c->stream, c->ev_out, 0);   // the previous transfer still reads the snapshot
    cudaMemcpyAsync(c->out_y, c->y, nb, cudaMemcpyDeviceToDevice, c->stream);
    cudaMemcpyAsync(c->out_yp, c->yp, nb, cudaMemcpyDeviceToDevice, c->stream);
    cudaEventRecord(c->ev_snap, c->stream);
    cudaStreamWaitEvent(c->out_stream, c->ev_snap, 0);
    cudaMemcpyAsync(y, c->out_y, nb, cudaMemcpyDeviceToHost, c->out_stream);
    cudaMemcpyAsync(yprime, c->out_yp, nb, cudaMemcpyDeviceToHost, c->out_stream);
    cudaEventRecord(c->ev_out, c->out_stream);
    c->out_pending = 1;
    return true;
}

extern "C" int dkb_output_wait(void)
{
    DkbCtx *c = g_ctx;
    if (!c) return dkb_fail(DKB_E_STATE, "dkb_init not called");
    if (c->out_pending) {
        CUDA_TRY(cudaEventSynchronize(c->ev_out));
        c->out_pending = 0;
    }
    return 0;
}

static void copy_out(DkbCtx *c, double *y, double *yprime, bool may_defer = false)
{
    const size_t nb = sizeof(double) * c->neq;
    const cudaMemcpyKind kind = c->opt_device_io ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (may_defer && c->opt_async_out && !c->opt_device_io && y && yprime && copy_out_async(c, y, yprime)) {
        cudaStreamSynchronize(c->stream);   // the step itself is finished (counters, t are final); only the transfer is in flight
        dkb_cuda_check("copy_out");
        return;
    }
    if (c->out_pending) {                   // an earlier deferred transfer targets the same arrays: let it finish first
        cudaEventSynchronize(c->ev_out);
        c->out_pending = 0;
    }
    if (y && y != c->y) cudaMemcpyAsync(y, c->y, nb, kind, c->stream);
    if (yprime && yprime != c->yp) cudaMemcpyAsync(yprime, c->yp, nb, kind, c->stream);
    cudaStreamSynchronize(c->stream);
    dkb_cuda_check("copy_out");
}

// ---------------------------------------------------------------- nonlinear solvers
struct Cb {
    dkb_res_t res;
    dkb_jack_t jac;
    dkb_psol_t psol;
    double *rpar;
    int *ipar;
};

static void call_res(DkbCtx *c, const Cb &cb, double t, const double *y, const double *ydot, double cj,
                     double *delta, int *ires)
{
    ProfScope ps(PROF_RES);
    cb.res(&t, y, ydot, &cj, delta, ires, cb.rpar, cb.ipar);
    (void)c;
}

// dnsk, src/daskr_new.f90:2881-3034
static void k_dnsk(DkbCtx *c, double t, const Cb &cb, int *iwm, double cj, double sqrtn, double rsqrtn,
                   double epslin, double epscon, double *s, double confac, double tolnew, int muldel, int maxit,
                   int *ires, int *iersl, int *iernew)
{
    int *IWM = iwm - 1;
    const i64 neq = c->neq;
    int m = 0;
    double deltanorm, oldnorm = 0.0, rate, rhok;
    bool converged = false;

    v_zero(neq, c->e);
    for (;;) {
        IWM[LNNI] = IWM[LNNI] + 1;
        if (muldel == 1) v_scale(neq, confac, c->delta);
        v_copy(neq, c->delta, c->savr);

        k_dslvk(neq, c->y, t, c->yp, c->savr, c->delta, c->wt, iwm, cb.res, ires, cb.psol, iersl, cj, epslin,
                sqrtn, rsqrtn, &rhok, cb.rpar, cb.ipar);
        if ((*ires != 0) || (*iersl != 0)) break;

        // y -= delta; e -= delta; ydot -= cj*delta  (:2988-2990), then ||delta|| (:2993)
        {
            ProfScope ps(PROF_VEC);
            v_newton_update(neq, c->delta, cj, c->y, c->e, c->yp);
        }
        deltanorm = r_wrms(neq, c->delta, c->wt);
        if (m == 0) {
            oldnorm = deltanorm;
            if (deltanorm <= tolnew) {
                converged = true;
                break;
            }
        } else {
            rate = pow(deltanorm / oldnorm, 1.0 / m);
            if (rate > 0.9) break;
            *s = rate / (1.0 - rate);
        }
        if ((*s) * deltanorm <= epscon) {
            converged = true;
            break;
        }
        m = m + 1;
        if (m >= maxit) break;

        *ires = 0;
        call_res(c, cb, t, c->y, c->yp, cj, c->delta, ires);
        IWM[LNRE] = IWM[LNRE] + 1;
        if (*ires < 0) break;
    }
    if (converged)
        *iernew = 0;
    else if ((*ires <= -2) || (*iersl < 0))
        *iernew = -1;
    else
        *iernew = 1;
}

// dnedk, src/daskr_new.f90:2660-2879
static void k_dnedk(DkbCtx *c, double t, const Cb &cb, double h, int jstart, int *idid, const double *gama,
                    int *iwm, double cj, double *cjold, double cjlast, double *s, double epli, double sqrtn,
                    double rsqrtn, double epscon, int *jcalc, int jflag, int kp1, int nonneg, int ntype,
                    int *iernls)
{
    const int muldel = 0, maxit = 4;
    const double xrate = 0.25;
    int *IWM = iwm - 1;
    const i64 neq = c->neq;
    const int neq32 = (int)neq;
    int iernew = 0, ierpj = 0, iersl = 0, iertyp = 0, ires = 0;
    double deltanorm, epslin, temp1 = 0.0, temp2, tolnew;
    double cc[6], dd[6];

    if (ntype != 1) {
        iertyp = 1;
        goto L380;
    }
    if (jstart == 0) {
        *cjold = cj;
        *jcalc = -1;
        *s = 1e2;
    }
    *iernls = 0;
    if (jflag != 0) {
        temp1 = (1.0 - xrate) / (1.0 + xrate);
        temp2 = 1 / temp1;
        if ((cj / (*cjold) < temp1) || (cj / (*cjold) > temp2)) *jcalc = -1;
        if (cj != cjlast) *s = 1e2;
    } else {
        *jcalc = 0;
    }

L300:
    ierpj = 0;
    ires = 0;
    iersl = 0;
    iernew = 0;

    // predictor (:2808-2813): y = sum phi_j, ydot = sum gama_j phi_j
    for (int j = 0; j < 6; ++j) {
        cc[j] = 1.0;
        dd[j] = (j >= 1 && j < kp1) ? gama[j] : 0.0;
    }
    {
        ProfScope ps(PROF_VEC);
        v_predict(neq, kp1, c->phi, cc, dd, c->y, c->yp);
    }

    epslin = epli * epscon;
    tolnew = epslin;

    call_res(c, cb, t, c->y, c->yp, cj, c->delta, &ires);
    IWM[LNRE] = IWM[LNRE] + 1;
    if (ires < 0) goto L380;

    if (*jcalc == -1) {
        *jcalc = 0;
        {
            ProfScope ps(PROF_JAC);
            cb.jac(cb.res, &ires, &neq32, &t, c->y, c->yp, c->wt, c->delta, c->e, &h, &cj, c->wp, c->iwp, &ierpj,
                   cb.rpar, cb.ipar);
        }
        IWM[LNJE] = IWM[LNJE] + 1;
        *cjold = cj;
        *s = 1e2;
        if (ires < 0) goto L380;
        if (ierpj != 0) goto L380;
    }

    k_dnsk(c, t, cb, iwm, cj, sqrtn, rsqrtn, epslin, epscon, s, temp1, tolnew, muldel, maxit, &ires, &iersl,
           &iernew);

    if (iernew > 0 && *jcalc != 0) {
        *jcalc = -1;
        goto L300;
    }
    if (iernew != 0) goto L380;

    if (nonneg == 0) goto L390;
    v_min0(neq, c->y, c->delta);
    deltanorm = r_wrms(neq, c->delta, c->wt);
    if (deltanorm > epscon) goto L380;
    v_lin2(neq, 1.0, c->e, -1.0, c->delta, c->e);
    goto L390;

L380:
    if ((ires <= -2) || (iersl < 0) || (iertyp != 0)) {
        *iernls = -1;
        if (ires <= -2) *idid = -11;
        if (iersl < 0) *idid = -13;
        if (iertyp != 0) *idid = -15;
    } else {
        *iernls = 1;
        if (ires == -1) *idid = -10;
        if (ierpj != 0) *idid = -5;
        if (iersl > 0) *idid = -14;
    }
L390:
    *jcalc = 1;
}

// ddstp, src/daskr_new.f90:149-567
static void k_ddstp(DkbCtx *c, double *t, const Cb &cb, double *h, int *jstart, int *idid, int *iwm,
                    double *alpha0_, double *beta0_, double *gama0_, double *psi0_, double *sigma0_, double *cj,
                    double *cjold, double *hold, double *s, double hmin, double epli, double sqrtn, double rsqrtn,
                    double epscon, int *iphase, int *jcalc, int jflg, int *k, int *kold, int *ns, int nonneg,
                    int ntype)
{
    double nrm3[3];   // error estimates of BLOCK 4, one fused pass
    int nnorm;
    bool fused3;
    int *IWM = iwm - 1;
    const i64 neq = c->neq;
    double *alpha = alpha0_ - 1, *beta = beta0_ - 1, *gama = gama0_ - 1, *psi = psi0_ - 1, *sigma = sigma0_ - 1;
#define PHI_(j) c->phi[(j) - 1]
    int iernls = 0, kdiff, km1, knew = 0, kp1, kp2, ncf, nef, nsp1;
    double alpha0, alphas, cjlast, ck, enorm, erk, erkm1 = 0.0, erkm2, erkp1, err, est = 0.0, hnew, r = 0.0, temp1,
           temp2, terk, terkm1 = 0.0, terkm2, terkp1, told;

    told = *t;
    ncf = 0;
    nef = 0;
    if (*jstart == 0) {
        *k = 1;
        *kold = 0;
        *hold = 0.0;
        psi[1] = *h;
        *cj = 1 / (*h);
        *iphase = 0;
        *ns = 0;
    }

    for (;;) {
        kp1 = *k + 1;
        kp2 = *k + 2;
        km1 = *k - 1;
        if ((*h != *hold) || (*k != *kold)) *ns = 0;
        *ns = std::min(*ns + 1, *kold + 2);
        nsp1 = *ns + 1;

        if (kp1 >= *ns) {
            beta[1] = 1.0;
            alpha[1] = 1.0;
            temp1 = *h;
            gama[1] = 0.0;
            sigma[1] = 1.0;
            for (int i = 2; i <= kp1; ++i) {
                temp2 = psi[i - 1];
                psi[i - 1] = temp1;
                beta[i] = beta[i - 1] * psi[i - 1] / temp2;
                temp1 = temp2 + *h;
                alpha[i] = *h / temp1;
                sigma[i] = (i - 1) * sigma[i - 1] * alpha[i];
                gama[i] = gama[i - 1] + alpha[i - 1] / (*h);
            }
            psi[kp1] = temp1;
        }
        alphas = 0.0;
        alpha0 = 0.0;
        for (int i = 1; i <= *k; ++i) {
            alphas = alphas - 1.0 / i;
            alpha0 = alpha0 - alpha[i];
        }
        cjlast = *cj;
        *cj = -alphas / (*h);
        ck = fabs(alpha[kp1] + alphas - alpha0);
        ck = std::max(ck, alpha[kp1]);

        // phi -> phi*  (:333-337)
        if (kp1 >= nsp1) {
            ProfScope ps(PROF_VEC);
            for (int j = nsp1; j <= kp1; ++j) v_scale(neq, beta[j], PHI_(j));
        }
        *t = *t + *h;
        *idid = 1;

        k_dnedk(c, *t, cb, *h, *jstart, idid, gama0_, iwm, *cj, cjold, cjlast, s, epli, sqrtn, rsqrtn, epscon, jcalc,
                jflg, kp1, nonneg, ntype, &iernls);
        if (iernls != 0) goto L600;

        // BLOCK 4: error estimates at orders k, k-1, k-2 (:360-378)
        // The three norms of this block in one pass over e, vt, phi(:,k+1), phi(:,k) (`delta` is scratch here in the
        // reference and is not read again before the next Newton iteration overwrites it).
        nnorm = (*k == 1) ? 1 : (*k == 2) ? 2 : 3;
        fused3 = r_wrms3(neq, nnorm, c->e, c->vt, PHI_(kp1), PHI_(*k), nrm3);
        enorm = fused3 ? nrm3[0] : r_wrms(neq, c->e, c->vt);
        erk = sigma[*k + 1] * enorm;
        terk = (*k + 1) * erk;
        est = erk;
        knew = *k;

        if (*k == 1) goto L430;
        if (!fused3) v_lin2(neq, 1.0, PHI_(kp1), 1.0, c->e, c->delta);
        erkm1 = sigma[*k] * (fused3 ? nrm3[1] : r_wrms(neq, c->delta, c->vt));
        terkm1 = (*k) * erkm1;
        if (*k > 2) goto L410;
        if (terkm1 <= terk / 2) goto L420;
        goto L430;
    L410:
        if (!fused3) v_lin2(neq, 1.0, c->delta, 1.0, PHI_(*k), c->delta);
        erkm2 = sigma[*k - 1] * (fused3 ? nrm3[2] : r_wrms(neq, c->delta, c->vt));
        terkm2 = (*k - 1) * erkm2;
        if (std::max(terkm1, terkm2) > terk) goto L430;
    L420:
        knew = *k - 1;
        est = erkm1;
    L430:
        err = ck * enorm;
        if (err > 1.0) goto L600;

        // BLOCK 5: successful step
        *idid = 1;
        IWM[LNST] = IWM[LNST] + 1;
        kdiff = *k - *kold;
        *kold = *k;
        *hold = *h;

        if (knew == km1 || *k == IWM[LMXORD]) *iphase = 1;
        if (*iphase == 0) goto L545;
        if (knew == km1) goto L540;
        if (*k == IWM[LMXORD]) goto L550;
        if (kp1 >= *ns || kdiff == 1) goto L550;

        v_lin2(neq, 1.0, c->e, -1.0, PHI_(kp2), c->delta);
        erkp1 = r_wrms(neq, c->delta, c->vt) / (*k + 2);
        terkp1 = (*k + 2) * erkp1;
        if (*k > 1) goto L520;
        if (terkp1 >= terk / 2) goto L550;
        goto L530;
    L520:
        if (terkm1 <= std::min(terk, terkp1)) goto L540;
        if (terkp1 >= terk || *k == IWM[LMXORD]) goto L550;
    L530:
        *k = kp1;
        est = erkp1;
        goto L550;
    L540:
        *k = km1;
        est = erkm1;
        goto L550;
    L545:
        *k = kp1;
        hnew = 2 * (*h);
        *h = hnew;
        goto L575;
    L550:
        hnew = *h;
        temp2 = (double)(*k + 1);
        r = pow(2 * est + 1e-4, -1 / temp2);
        if (r < 2.0) goto L555;
        hnew = 2 * (*h);
        goto L560;
    L555:
        if (r > 1.0) goto L560;
        r = std::max(0.5, std::min(0.9, r));
        hnew = (*h) * r;
    L560:
        *h = hnew;
    L575: {
        // phi update cascade (:458-466) in one multi-vector kernel
        ProfScope ps(PROF_VEC);
        v_phi_update(neq, kp1, (*kold != IWM[LMXORD]) ? 1 : 0, c->phi, c->e);
    }
        *jstart = 1;
        return;

        // BLOCK 6: unsuccessful step
    L600:
        *iphase = 1;
        *t = told;
        if (kp1 >= nsp1)
            for (int j = nsp1; j <= kp1; ++j) {
                temp1 = 1 / beta[j];
                v_scale(neq, temp1, PHI_(j));
            }
        for (int i = 2; i <= kp1; ++i) psi[i - 1] = psi[i] - *h;

        if (iernls == 0) goto L660;
        IWM[LNCFN] = IWM[LNCFN] + 1;
        if (iernls < 0) goto L675;
        ncf = ncf + 1;
        r = 0.25;
        *h = (*h) * r;
        if (ncf < 10 && fabs(*h) >= hmin) goto L690;
        if (*idid == 1) *idid = -7;
        if (nef >= 3) *idid = -9;
        goto L675;
    L660:
        nef = nef + 1;
        IWM[LETF] = IWM[LETF] + 1;
        if (nef > 1) goto L665;
        *k = knew;
        temp2 = (double)(*k + 1);
        r = 0.9 * pow(2 * est + 1e-4, -1 / temp2);
        r = std::max(0.25, std::min(0.9, r));
        *h = (*h) * r;
        if (fabs(*h) >= hmin) goto L690;
        *idid = -6;
        goto L675;
    L665:
        if (nef > 2) goto L670;
        *k = knew;
        r = 0.25;
        *h = r * (*h);
        if (fabs(*h) >= hmin) goto L690;
        *idid = -6;
        goto L675;
    L670:
        *k = 1;
        r = 0.25;
        *h = r * (*h);
        if (fabs(*h) >= hmin) goto L690;
        *idid = -6;
        goto L675;
    L675:
        interp(c, *t, *t, *k, psi);
        *jstart = 1;
        if (*idid >= 0) *idid = -7;
        return;
    L690:
        if (*kold == 0) {
            psi[1] = *h;
            v_scale(neq, r, PHI_(2));
        }
    }
}

// ---------------------------------------------------------------- initial-condition path
// dfnrmk, src/daskr_new.f90:2568-2658
static void k_dfnrmk(DkbCtx *c, const Cb &cb, const double *y, double t, const double *ydot, double *savr,
                     double *r, double cj, double tscale, double sqrtn, double rsqrtn, int *ires, int irin,
                     int *ierr, double *rnorm, double epslin, double *wk)
{
    (void)sqrtn;
    const i64 neq = c->neq;
    const int neq32 = (int)neq;
    if (irin == 0) {
        *ires = 0;
        call_res(c, cb, t, y, ydot, cj, savr, ires);
        if (*ires < 0) return;
    }
    v_copy(neq, savr, r);
    v_scale_to(neq, rsqrtn, c->wt, c->wght);   // wght = wt/sqrt(N) (:2647); wt itself is left alone
    *ierr = 0;
    cb.psol(&neq32, &t, y, ydot, savr, wk, &cj, c->wght, c->wp, c->iwp, r, &epslin, ierr, cb.rpar, cb.ipar);
    if (*ierr != 0) return;
    *rnorm = r_wrms(neq, r, c->wt);
    if (tscale > 0.0) *rnorm = (*rnorm) * tscale * fabs(cj);
}

// dlinsk, src/daskr_new.f90:2365-2566 (constraint checking not supported on device yet)
static void k_dlinsk(DkbCtx *c, const Cb &cb, double t, double cj, double tscale, double *pnorm, double sqrtn,
                     double rsqrtn, int lsoff, double stptol, int *iret, int *ires, int *iwm, double *fnorm,
                     int icopt, double epslin)
{
    const double alpha = 1e-4;
    int *IWM = iwm - 1;
    const i64 neq = c->neq;
    int ierr = 0;
    double f1norm, f1normp, fnormp = 0.0, ratio, rl, rlmin, slpi;
    double *p = c->delta, *r = c->e, *ynew = c->yic, *ydotnew = c->ypic, *pwk = c->pwk;

    f1norm = ((*fnorm) * (*fnorm)) / 2;
    ratio = 1.0;
    rl = 1.0;
    // constraint violations rescale the step until there is none (:2486-2504); rlx = 0.4 (dnsik:2276)
    if (c->s_icnflg != 0) {
        const double rlx = 0.4, fac = 0.6, fac2 = 0.9;
        double tau = *pnorm;
        for (;;) {
            v_dyypnw(neq, c->y, c->yp, cj, rl, p, icopt, c->d_id, ynew, ydotnew);
            double rr[2];
            r_cnstr(neq, c->y, ynew, c->d_icnstr, 46);
            scal_fetch(46, 2, rr);
            if (rr[1] > 0.0) tau = fac * tau;                          // a hard violation (dcnstr:611-649)
            else if (rr[0] >= rlx) tau = fac2 * tau * rlx / rr[0];     // too large a relative change (:654-657)
            else break;
            const double ratio1 = tau / (*pnorm);
            ratio = ratio * ratio1;
            v_scale(neq, ratio1, p);
            *pnorm = tau;
            if (*pnorm <= stptol) {
                *iret = 1;
                return;
            }
        }
    }
    slpi = -2 * f1norm * ratio;
    rlmin = stptol / (*pnorm);

    for (;;) {
        v_dyypnw(neq, c->y, c->yp, cj, rl, p, icopt, c->d_id, ynew, ydotnew);
        k_dfnrmk(c, cb, ynew, t, ydotnew, c->savr, r, cj, tscale, sqrtn, rsqrtn, ires, 0, &ierr, &fnormp, epslin,
                 pwk);
        IWM[LNRE] = IWM[LNRE] + 1;
        if (*ires >= 0) IWM[LNPS] = IWM[LNPS] + 1;
        if (*ires != 0 || ierr != 0) {
            *iret = 2;
            return;
        }
        if (lsoff != 1) {
            f1normp = (fnormp * fnormp) / 2;
            if (f1normp > f1norm + alpha * slpi * rl) {
                if (rl < rlmin) {
                    *iret = 1;
                    return;
                }
                rl = rl / 2;
                continue;
            }
        }
        *iret = 0;
        v_copy(neq, ynew, c->y);
        v_copy(neq, ydotnew, c->yp);
        *fnorm = fnormp;
        return;
    }
}

// dnsik, src/daskr_new.f90:2175-2363
static void k_dnsik(DkbCtx *c, const Cb &cb, double t, int icopt, int *iwm, double cj, double tscale, double sqrtn,
                    double rsqrtn, double epslin, double epscon, double ratemax, int maxit, double stptol,
                    int *iernew)
{
    int *IWM = iwm - 1;
    const i64 neq = c->neq;
    int ierr = 0, iersl = 0, ires = 0, iret = 0, m;
    double deltanorm, fnorm = 0.0, fnorm0, oldfnorm, rate, rhok = 0.0;
    bool iserror, ismaxit;
    const int lsoff = IWM[LLSOFF];
    rate = 1.0;
    *iernew = 0;

    v_copy(neq, c->delta, c->savr);
    k_dfnrmk(c, cb, c->y, t, c->yp, c->savr, c->e, cj, tscale, sqrtn, rsqrtn, &ires, 1, &ierr, &fnorm, epslin,
             c->pwk);
    IWM[LNPS] = IWM[LNPS] + 1;
    if (ierr != 0) {
        *iernew = 3;
        return;
    }
    if (fnorm <= epscon) return;

    fnorm0 = fnorm;
    iserror = false;
    ismaxit = false;
    m = 0;
    for (;;) {
        IWM[LNNI] = IWM[LNNI] + 1;
        k_dslvk(neq, c->y, t, c->yp, c->savr, c->delta, c->wt, iwm, cb.res, &ires, cb.psol, &iersl, cj, epslin,
                sqrtn, rsqrtn, &rhok, cb.rpar, cb.ipar);
        if ((ires != 0) || (iersl != 0)) {
            iserror = true;
            break;
        }
        deltanorm = r_wrms(neq, c->delta, c->wt);
        if (deltanorm == 0.0) break;

        oldfnorm = fnorm;
        k_dlinsk(c, cb, t, cj, tscale, &deltanorm, sqrtn, rsqrtn, lsoff, stptol, &iret, &ires, iwm, &fnorm, icopt,
                 epslin);
        rate = fnorm / oldfnorm;
        if (iret != 0) {
            iserror = true;
            break;
        }
        if (fnorm <= epscon) break;
        m = m + 1;
        if (m >= maxit) {
            ismaxit = true;
            break;
        }
        v_copy(neq, c->savr, c->delta);
    }
    if (ismaxit) {
        if ((rate <= ratemax) || (fnorm <= fnorm0 / 10))
            *iernew = 1;
        else
            *iernew = 2;
    }
    if (iserror) {
        if ((ires <= -2) || (iersl < 0)) {
            *iernew = -1;
        } else {
            *iernew = 3;
            if ((ires == 0) && (iersl == 1) && (m >= 2) && (rate < 1.0)) *iernew = 1;
        }
    }
}

// ddasik, src/daskr_new.f90:2004-2173
static void k_ddasik(DkbCtx *c, const Cb &cb, double t, int icopt, double h, double tscale, int *jskip, int *iwm,
                     double cj, double epli, double sqrtn, double rsqrtn, double epscon, double ratemax,
                     double stptol, int jflg, int *iernls)
{
    int *IWM = iwm - 1;
    const i64 neq = c->neq;
    const int neq32 = (int)neq;
    int iernew = 0, ierpj = 0, ires, nj = 0;
    const int maxnit = IWM[LMXNIT], maxnj = IWM[LMXNJ];
    const double epslin = epli * epscon;
    bool failed = false;

    ires = 0;
    call_res(c, cb, t, c->y, c->yp, cj, c->delta, &ires);
    IWM[LNRE] = IWM[LNRE] + 1;
    if (ires < 0) failed = true;

    if (!failed) {
        for (;;) {
            *iernls = 0;
            ierpj = 0;
            ires = 0;
            iernew = 0;
            if ((jflg == 1) && (*jskip == 0)) {
                nj = nj + 1;
                IWM[LNJE] = IWM[LNJE] + 1;
                {
                    ProfScope ps(PROF_JAC);
                    cb.jac(cb.res, &ires, &neq32, &t, c->y, c->yp, c->wt, c->delta, c->e, &h, &cj, c->wp, c->iwp,
                           &ierpj, cb.rpar, cb.ipar);
                }
                if ((ires < 0) || (ierpj != 0)) {
                    failed = true;
                    break;
                }
            }
            *jskip = 0;
            k_dnsik(c, cb, t, icopt, iwm, cj, tscale, sqrtn, rsqrtn, epslin, epscon, ratemax, maxnit, stptol,
                    &iernew);
            if ((iernew == 1) && (nj < maxnj) && (jflg == 1)) {
                v_copy(neq, c->savr, c->delta);
                continue;
            }
            break;
        }
    }
    if (failed) {
        *iernls = 2;
        if (ires <= -2) *iernls = -1;
        return;
    }
    if (iernew != 0) *iernls = std::min(iernew, 2);
}

// DDASIC, src/daskr.f:2911-3077
static void k_ddasic(DkbCtx *c, const Cb &cb, double x, int icopt, double *h, double tscale, int nic, int *idid,
                     int *iwm, double epli, double sqrtn, double rsqrtn, double epconi, double stptol, int jflg)
{
    const double rhcut = 0.1, ratemx = 0.8;
    int *IWM = iwm - 1;
    const i64 neq = c->neq;
    const int mxnh = IWM[LMXNH];
    int nh = 1, jskip = 0, iernls = 0;
    double cj;
    *idid = 1;
    if (nic == 2) jskip = 1;
    v_copy(neq, c->y, c->phi[0]);
    v_copy(neq, c->yp, c->phi[1]);
    cj = (icopt == 2) ? 0.0 : 1.0 / (*h);

    for (;;) {
        k_ddasik(c, cb, x, icopt, *h, tscale, &jskip, iwm, cj, epli, sqrtn, rsqrtn, epconi, ratemx, stptol, jflg,
                 &iernls);
        if (iernls == 0) return;
        IWM[LNCFN] = IWM[LNCFN] + 1;
        jskip = 0;
        if (iernls == -1) break;
        if (icopt == 2) break;
        if (nh == mxnh) break;
        nh = nh + 1;
        *h = (*h) * rhcut;
        cj = 1.0 / (*h);
        if (iernls == 1) continue;
        v_copy(neq, c->phi[0], c->y);
        v_copy(neq, c->phi[1], c->yp);
    }
    *idid = -12;
}

// ---------------------------------------------------------------- state export / import
// The reference keeps its whole state in the caller's rwork / iwork, so a run can be checkpointed by saving those
// arrays (plus y, yprime) and continued later with info(1) = 1.  Here the length-NEQ segments of rwork live on the
// device; these three calls move exactly what a continuation call reads -- y, yprime, the PHI array, the
// preconditioner blocks WP / IWP, the ID array and vector tolerances, and the SAVEd scalars -- to and from one host
// buffer.  rwork(1:60+3*nrt) and iwork(1:40+...) stay with the caller, as in the reference.
struct DkbStateHdr {
    char magic[8];
    i64 neq, neq_g, lenwp, leniwp, s_lid;
    int maxl, kmp, has_id, has_tol, s_lenid, s_nonneg, s_ncphi, s_icnflg;
};

static i64 state_bytes(const DkbCtx *c)
{
    i64 n = (i64)sizeof(DkbStateHdr) + (i64)sizeof(double) * c->neq * 8 + (i64)sizeof(double) * c->lenwp +
            (i64)sizeof(int) * c->leniwp;
    if (c->d_id) n += (i64)sizeof(int) * c->neq;
    if (c->d_rtol) n += (i64)sizeof(double) * c->neq * 2;
    return n;
}

extern "C" int64_t dkb_state_bytes(void)
{
    DkbCtx *c = g_ctx;
    if (!c || !c->arena) return 0;
    return state_bytes(c);
}

extern "C" int dkb_state_export(void *buf, int64_t nbytes)
{
    DkbCtx *c = g_ctx;
    if (!c || !c->arena) return dkb_fail(DKB_E_STATE, "dkb_state_export: no solver state (call dkb_daskr first)");
    if (nbytes < state_bytes(c)) return dkb_fail(DKB_E_ARG, "dkb_state_export: buffer too small (need %lld bytes)", (long long)state_bytes(c));
    DkbStateHdr h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "DKBSTAT1", 8);
    h.neq = c->neq; h.neq_g = c->neq_g; h.lenwp = c->lenwp; h.leniwp = c->leniwp; h.s_lid = c->s_lid;
    h.maxl = c->maxl; h.kmp = c->kmp; h.has_id = c->d_id != nullptr; h.has_tol = c->d_rtol != nullptr;
    h.s_lenid = c->s_lenid; h.s_nonneg = c->s_nonneg; h.s_ncphi = c->s_ncphi; h.s_icnflg = c->s_icnflg;
    char *p = (char *)buf;
    memcpy(p, &h, sizeof h);
    p += sizeof h;
    auto out = [&](const void *d, size_t n) {
        cudaMemcpyAsync(p, d, n, cudaMemcpyDeviceToHost, c->stream);
        p += n;
    };
    const size_t nb = sizeof(double) * c->neq;
    out(c->y, nb);
    out(c->yp, nb);
    for (int j = 0; j < 6; ++j) out(c->phi[j], nb);
    out(c->wp, sizeof(double) * c->lenwp);
    out(c->iwp, sizeof(int) * c->leniwp);
    if (c->d_id) out(c->d_id, sizeof(int) * c->neq);
    if (c->d_rtol) { out(c->d_rtol, nb); out(c->d_atol, nb); }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// After dkb_init, dkb_setup_rbdpre and the problem's setpar (same mesh, same ranks).  The next dkb_daskr call must
// be a continuation call (info(1) = 1) with the rwork / iwork saved next to this buffer.
extern "C" int dkb_state_import(const void *buf, int64_t nbytes)
{
    DkbCtx *c = g_ctx;
    if (!c) return dkb_fail(DKB_E_STATE, "dkb_init not called");
    DkbStateHdr h;
    if (nbytes < (int64_t)sizeof h) return dkb_fail(DKB_E_ARG, "dkb_state_import: buffer too small");
    memcpy(&h, buf, sizeof h);
    if (memcmp(h.magic, "DKBSTAT1", 8) != 0) return dkb_fail(DKB_E_ARG, "dkb_state_import: not a daskr_b200 state buffer");
    if (c->rbd_set && h.neq != (i64)c->ns * c->npts)
        return dkb_fail(DKB_E_ARG, "dkb_state_import: state has neq = %lld, this mesh slab has %lld", (long long)h.neq,
                        (long long)((i64)c->ns * c->npts));
    const i64 rbd_wp = c->rbg_set ? (i64)c->ngx * c->ngy * c->ns * c->ns : c->rbd_set ? tpp_lut_doubles(c->ns, c->npts) : 0;
    const i64 rbd_iwp = c->rbg_set ? (i64)c->ngx * c->ngy * c->ns : c->rbd_set ? tpp_pt_ints(c->ns, c->npts) : 0;
    if (alloc_solver(c, h.neq, h.maxl, h.kmp, 0, h.has_id, 0, h.has_tol, h.lenwp - rbd_wp, h.leniwp - rbd_iwp) != 0) return DKB_E_CUDA;
    if (c->lenwp != h.lenwp || c->leniwp != h.leniwp) return dkb_fail(DKB_E_ARG, "dkb_state_import: preconditioner storage differs");
    c->neq_g = h.neq_g; c->s_lid = h.s_lid; c->s_lenid = h.s_lenid; c->s_nonneg = h.s_nonneg; c->s_ncphi = h.s_ncphi;
    c->s_icnflg = h.s_icnflg;
    i64 need = state_bytes(c);
    if (nbytes < need) return dkb_fail(DKB_E_ARG, "dkb_state_import: buffer truncated (need %lld bytes)", (long long)need);
    const char *p = (const char *)buf + sizeof h;
    auto in = [&](void *d, size_t n) {
        cudaMemcpyAsync(d, p, n, cudaMemcpyHostToDevice, c->stream);
        p += n;
    };
    const size_t nb = sizeof(double) * c->neq;
    in(c->y, nb);
    in(c->yp, nb);
    for (int j = 0; j < 6; ++j) in(c->phi[j], nb);
    in(c->wp, sizeof(double) * c->lenwp);
    in(c->iwp, sizeof(int) * c->leniwp);
    if (c->d_id) in(c->d_id, sizeof(int) * c->neq);
    if (c->d_rtol) { in(c->d_rtol, nb); in(c->d_atol, nb); }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---------------------------------------------------------------- rootfinding (DRCHEK / DROOTS)
static inline double sign1(double x) { return signbit(x) ? -1.0 : 1.0; }   // SIGN(1.0D0, x)

// DROOTS, src/daskr.f:2676-2910: scalar logic on the nrt root-function values (host).
// SAVEd between calls: alpha, x2, imax, last.
static void h_droots(int nrt, double hmin, int *jflag, double *x0, double *x1, double *r0, double *r1, double *rx,
                     double *x, int *jroot)
{
    static double alpha, x2;
    static int imax, last;
    int imxold, nxlast = 0;
    double t2, tmax, fracint, fracsub;
    bool zroot, sgnchg, xroot = false;

    if (*jflag == 1) goto L200;
    imax = 0;
    tmax = 0.0;
    zroot = false;
    for (int i = 0; i < nrt; ++i) {
        if (!(fabs(r1[i]) > 0.0)) { zroot = true; continue; }
        if (sign1(r0[i]) == sign1(r1[i])) continue;
        t2 = fabs(r1[i] / (r1[i] - r0[i]));
        if (t2 <= tmax) continue;
        tmax = t2;
        imax = i + 1;
    }
    sgnchg = imax > 0;
    if (!sgnchg) goto L400;
    xroot = false;
    nxlast = 0;
    last = 1;
L150:
    if (xroot) goto L300;
    if (nxlast != last) alpha = 1.0;
    else if (last != 0) alpha = 0.5 * alpha;
    else alpha = 2.0 * alpha;
    x2 = *x1 - (*x1 - *x0) * r1[imax - 1] / (r1[imax - 1] - alpha * r0[imax - 1]);
    if (fabs(x2 - *x0) < 0.5 * hmin) {
        fracint = fabs(*x1 - *x0) / hmin;
        fracsub = (fracint > 5.0) ? 0.1 : 0.5 / fracint;
        x2 = *x0 + fracsub * (*x1 - *x0);
    }
    if (fabs(*x1 - x2) < 0.5 * hmin) {
        fracint = fabs(*x1 - *x0) / hmin;
        fracsub = (fracint > 5.0) ? 0.1 : 0.5 / fracint;
        x2 = *x1 - fracsub * (*x1 - *x0);
    }
    *jflag = 1;
    *x = x2;
    return;
L200:
    imxold = imax;
    imax = 0;
    tmax = 0.0;
    zroot = false;
    for (int i = 0; i < nrt; ++i) {
        if (!(fabs(rx[i]) > 0.0)) { zroot = true; continue; }
        if (sign1(r0[i]) == sign1(rx[i])) continue;
        t2 = fabs(rx[i] / (rx[i] - r0[i]));
        if (t2 <= tmax) continue;
        tmax = t2;
        imax = i + 1;
    }
    if (imax > 0) {
        sgnchg = true;
    } else {
        sgnchg = false;
        imax = imxold;
    }
    nxlast = last;
    if (sgnchg) {
        *x1 = x2;
        for (int i = 0; i < nrt; ++i) r1[i] = rx[i];
        last = 1;
        xroot = false;
    } else if (zroot) {
        *x1 = x2;
        for (int i = 0; i < nrt; ++i) r1[i] = rx[i];
        xroot = true;
    } else {
        for (int i = 0; i < nrt; ++i) r0[i] = rx[i];
        *x0 = x2;
        last = 0;
        xroot = false;
    }
    if (fabs(*x1 - *x0) <= hmin) xroot = true;
    goto L150;
L300:
    *jflag = 2;
    *x = *x1;
    for (int i = 0; i < nrt; ++i) rx[i] = r1[i];
    for (int i = 0; i < nrt; ++i) {
        jroot[i] = 0;
        if (fabs(r1[i]) == 0.0) {
            jroot[i] = (int)(-sign1(r0[i]));
            continue;
        }
        if (sign1(r0[i]) != sign1(r1[i])) jroot[i] = (int)sign1(r1[i] - r0[i]);
    }
    return;
L400:
    if (zroot) {
        *x = *x1;
        for (int i = 0; i < nrt; ++i) rx[i] = r1[i];
        for (int i = 0; i < nrt; ++i) {
            jroot[i] = 0;
            if (fabs(r1[i]) == 0.0) jroot[i] = (int)(-sign1(r0[i]));
        }
        *jflag = 3;
        return;
    }
    for (int i = 0; i < nrt; ++i) rx[i] = r1[i];
    *x = *x1;
    *jflag = 4;
}

// DRCHEK, src/daskr.f:2494-2675.  The interpolations (ddatrp) and the y += temp2*phi(:,2) update are
// device kernels on the internal y / yprime; rt sees device arrays and returns nrt host scalars.
static void k_drchek(DkbCtx *c, int job, dkb_rt_t rt, int nrt, double tn, double tout, const double *psi /*1-based*/,
                     int kold, double *r0, double *r1, double *rx, int *jroot, int *irt, double uround,
                     double *rwork, int *iwork, double *rpar, int *ipar)
{
    double *RW = rwork - 1;
    int *IW = iwork - 1;
    const i64 neq = c->neq;
    const int neq32 = (int)neq;
    const double h = psi[1];
    double t1, temp1, temp2, x = 0.0;
    bool zroot;
    int jflag;
    *irt = 0;
    for (int i = 0; i < nrt; ++i) jroot[i] = 0;
    const double hminr = (fabs(tn) + fabs(h)) * uround * 100.0;

    if (job == 1) {
        interp(c, tn, RW[LT0], kold, psi);
        rt(&neq32, &RW[LT0], c->y, c->yp, &nrt, r0, rpar, ipar);
        IW[LNRTE] = 1;
        zroot = false;
        for (int i = 0; i < nrt; ++i)
            if (fabs(r0[i]) == 0.0) zroot = true;
        if (!zroot) return;
        temp2 = std::max(hminr / fabs(h), 0.1);
        temp1 = temp2 * h;
        RW[LT0] = RW[LT0] + temp1;
        v_axpy(neq, temp2, c->phi[1], c->y);
        rt(&neq32, &RW[LT0], c->y, c->yp, &nrt, r0, rpar, ipar);
        IW[LNRTE] = IW[LNRTE] + 1;
        zroot = false;
        for (int i = 0; i < nrt; ++i)
            if (fabs(r0[i]) == 0.0) zroot = true;
        if (zroot) *irt = -1;
        return;
    }
    if (job == 2) {
        if (IW[LIRFND] != 0) {
            interp(c, tn, RW[LT0], kold, psi);
            rt(&neq32, &RW[LT0], c->y, c->yp, &nrt, r0, rpar, ipar);
            IW[LNRTE] = IW[LNRTE] + 1;
            zroot = false;
            for (int i = 0; i < nrt; ++i)
                if (fabs(r0[i]) == 0.0) {
                    zroot = true;
                    jroot[i] = 1;
                }
            if (zroot) {
                temp1 = (h >= 0.0) ? fabs(hminr) : -fabs(hminr);
                RW[LT0] = RW[LT0] + temp1;
                if ((RW[LT0] - tn) * h < 0.0) {
                    interp(c, tn, RW[LT0], kold, psi);
                } else {
                    temp2 = temp1 / h;
                    v_axpy(neq, temp2, c->phi[1], c->y);
                }
                rt(&neq32, &RW[LT0], c->y, c->yp, &nrt, r0, rpar, ipar);
                IW[LNRTE] = IW[LNRTE] + 1;
                for (int i = 0; i < nrt; ++i) {
                    if (fabs(r0[i]) > 0.0) continue;
                    if (jroot[i] == 1) {
                        *irt = -2;
                        return;
                    }
                    jroot[i] = (int)(-sign1(r0[i]));
                    *irt = 1;
                }
                if (*irt == 1) return;
            }
        }
        if (tn == RW[LTLAST]) return;
    }
    if ((tout - tn) * h >= 0.0) {
        t1 = tn;
    } else {
        t1 = tout;
        if ((t1 - RW[LT0]) * h <= 0.0) return;
    }
    interp(c, tn, t1, kold, psi);
    rt(&neq32, &t1, c->y, c->yp, &nrt, r1, rpar, ipar);
    IW[LNRTE] = IW[LNRTE] + 1;
    jflag = 0;
    for (;;) {
        h_droots(nrt, hminr, &jflag, &RW[LT0], &t1, r0, r1, rx, &x, jroot);
        if (jflag > 1) break;
        interp(c, tn, x, kold, psi);
        rt(&neq32, &x, c->y, c->yp, &nrt, rx, rpar, ipar);
        IW[LNRTE] = IW[LNRTE] + 1;
    }
    RW[LT0] = x;
    for (int i = 0; i < nrt; ++i) r0[i] = rx[i];
    if (jflag == 4) return;
    interp(c, tn, x, kold, psi);
    *irt = 1;
}

// ---------------------------------------------------------------- DASKR
extern "C" void dkb_daskr(dkb_res_t res, const int *neq_, double *t, double *y, double *yprime, const double *tout_,
                          int *info, double *rtol, double *atol, int *idid, double *rwork, const int64_t *lrw,
                          int *iwork, const int64_t *liw, double *rpar, int *ipar, dkb_jack_t jac, dkb_psol_t psol,
                          void *rt, const int *nrt_, int *jroot)
{
    dkb_rt_t rtf = (dkb_rt_t)rt;
    int irt = 0;
    DkbCtx *c = g_ctx;
    int *INFO = info - 1;
    if (!c) {
        fprintf(stderr, "daskr_b200: dkb_daskr called before dkb_init\n");
        INFO[1] = -1;
        *idid = -33;
        return;
    }
    const i64 neq = *neq_;
    const int nrt = *nrt_;
    const double tout = *tout_;
    int *IW = iwork - 1;
    double *RW = rwork - 1;
    Cb cb{res, jac, psol, rpar, ipar};
    Tol tol{INFO[2], rtol, atol};
    int i, mxord, index = 1, ier, nwt, nzflg;
    i64 lenic, lenrw, leniw, maxl;
    double tn = 0.0, h = 0.0, h0 = 0.0, hmin, uround, tdist, ypnorm, rh, tstop = 0.0, epconi, tscale, r, rtoli,
           atoli, tnext;
    bool done;
    const char *why = "";

    if (INFO[1] != 0) goto L100;

    // ---- initial call: input checks (src/daskr.f:1447-1661) ----
    for (i = 2; i <= 9; ++i)
        if (INFO[i] != 0 && INFO[i] != 1) { why = "info(2:9) out of range"; goto L701; }
    if (INFO[10] < 0 || INFO[10] > 3) { why = "info(10) out of range"; goto L701; }
    if (INFO[11] < 0 || INFO[11] > 2) { why = "info(11) out of range"; goto L701; }
    for (i = 12; i <= 17; ++i)
        if (INFO[i] != 0 && INFO[i] != 1) { why = "info(12:17) out of range"; goto L701; }
    if (INFO[18] < 0 || INFO[18] > 2) { why = "info(18) out of range"; goto L701; }
    if (neq <= 0) { why = "neq <= 0"; goto L701; }
    if (INFO[12] != 1) { why = "only the Krylov option info(12)=1 is implemented on device"; goto L701; }
    if (nrt < 0) { why = "nrt < 0"; goto L701; }
    if (nrt > 0 && rt == nullptr) { why = "nrt > 0 needs an rt routine (device arrays)"; goto L701; }
    if (c->rbd_set && neq != (i64)c->ns * c->npts) { why = "neq does not match setup_rbdpre slab (ns*mx*myloc)"; goto L701; }

    mxord = 5;
    if (INFO[9] != 0) {
        mxord = IW[LMXORD];
        if (mxord < 1 || mxord > 5) { why = "mxord out of range"; goto L701; }
    }
    IW[LMXORD] = mxord;

    // info(10): 0 none, 1 constraints in the IC calculation, 2 non-negativity in the integration, 3 both (daskr.f:1477-1494)
    c->s_icnflg = (INFO[10] == 1 || INFO[10] == 3) ? 1 : 0;
    c->s_nonneg = (INFO[10] == 2 || INFO[10] == 3) ? 1 : 0;
    c->s_lid = (c->s_icnflg != 0) ? LICNS + neq : LICNS;

    IW[LMITER] = INFO[12];
    if (INFO[13] == 0) {
        IW[LMAXL] = (int)std::min<i64>(5, neq);
        IW[LKMP] = IW[LMAXL];
        IW[LNRMAX] = 5;
        RW[LEPLI] = 0.05;
    } else {
        if (IW[LMAXL] < 1 || IW[LMAXL] > neq || IW[LMAXL] > DKB_MAXL_MAX) { why = "maxl out of range"; goto L701; }
        if (IW[LKMP] < 1 || IW[LKMP] > IW[LMAXL]) { why = "kmp out of range"; goto L701; }
        if (IW[LNRMAX] < 0) { why = "nrmax < 0"; goto L701; }
        if (RW[LEPLI] <= 0.0 || RW[LEPLI] >= 1.0) { why = "epli out of range"; goto L701; }
    }
    if (INFO[11] != 0) {
        if (INFO[17] == 0) {
            IW[LMXNIT] = 15;
            IW[LMXNJ] = 2;
            IW[LMXNH] = 5;
            IW[LLSOFF] = 0;
            RW[LEPIN] = 0.01;
        } else {
            if (IW[LMXNIT] <= 0 || IW[LMXNJ] <= 0 || IW[LMXNH] <= 0 || IW[LLSOFF] < 0 || IW[LLSOFF] > 1 ||
                RW[LEPIN] <= 0.0) { why = "IC controls out of range"; goto L701; }
        }
    }

    // work-array lengths: scalar headers only (the vector segments are device memory)
    lenic = (c->s_icnflg != 0) ? neq : 0;
    c->s_lenid = 0;
    if (INFO[11] == 1 || INFO[16] == 1) c->s_lenid = (int)neq;
    c->s_ncphi = mxord + 1;
    maxl = IW[LMAXL];
    lenrw = 60 + 3 * nrt;
    leniw = 40 + lenic + c->s_lenid;
    IW[LNIW] = (int)std::min<i64>(leniw, INT_MAX);
    IW[LNRW] = (int)lenrw;
    IW[LNPD] = 0;
    IW[LLOCWP] = 1;
    if (*lrw < lenrw) { why = "rwork too short (need 60 + 3*nrt)"; goto L701; }
    if (*liw < leniw) { why = "iwork too short (need 40 + lenic + lenid)"; goto L701; }

    if (alloc_solver(c, neq, (int)maxl, IW[LKMP], INFO[16], c->s_lenid > 0, lenic > 0, INFO[2], IW[LLNWP], IW[LLNIWP]) != 0) {
        why = c->err;
        goto L701;
    }
    c->neq_g = neq;
    if (c->nranks > 1) {
        double ng = (double)neq;
        cudaMemcpyAsync(c->d_scal + 50, &ng, sizeof(double), cudaMemcpyHostToDevice, c->stream);
        allreduce_sum(50, 1);
        scal_fetch(50, 1, &ng);
        c->neq_g = (i64)(ng + 0.5);
    }

    // ICNSTR legality (daskr.f:1608-1615); the consistency of y with it (dcnst0) is checked once y is on the device
    if (lenic > 0) {
        for (i64 ii = 1; ii <= neq; ++ii) {
            const int ici = IW[LICNS - 1 + ii];
            if (ici < -2 || ici > 2) { why = "ICNSTR entries must be in -2..2"; goto L701; }
        }
        cudaMemcpyAsync(c->d_icnstr, &IW[LICNS], sizeof(int) * neq, cudaMemcpyHostToDevice, c->stream);
    }
    index = 1;
    if (c->s_lenid > 0) {
        index = 0;
        for (i64 ii = 1; ii <= neq; ++ii) {
            int idi = IW[c->s_lid - 1 + ii];
            if (idi != 1 && idi != -1) { why = "ID array entries must be +1 / -1"; goto L701; }
            if (idi == -1) index = 1;
        }
        cudaMemcpyAsync(c->d_id, &IW[c->s_lid], sizeof(int) * neq, cudaMemcpyHostToDevice, c->stream);
    }
    if (tout == *t) { why = "tout == t"; goto L701; }
    if (INFO[7] != 0 && RW[LHMAX] <= 0.0) { why = "hmax <= 0"; goto L701; }

    IW[LNST] = 0;
    IW[LNRE] = 0;
    IW[LNJE] = 0;
    IW[LETF] = 0;
    IW[LNCFN] = 0;
    IW[LNNI] = 0;
    IW[LNLI] = 0;
    IW[LNPS] = 0;
    IW[LNCFL] = 0;
    IW[LNRTE] = 0;
    IW[LKPRIN] = INFO[18];
    *idid = 1;

    // y, yprime enter the device here (and only here: on continuation calls they are outputs)
    {
        const size_t nb = sizeof(double) * neq;
        const cudaMemcpyKind kind = c->opt_device_io ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        cudaMemcpyAsync(c->y, y, nb, kind, c->stream);
        cudaMemcpyAsync(c->yp, yprime, nb, kind, c->stream);
    }
    if (lenic > 0) {   // dcnst0 (daskr.f:1617-1622): the initial y must satisfy the constraints
        c->h_iflag[0] = INT_MAX;
        cudaMemcpyAsync(c->d_iflag, c->h_iflag, sizeof(int), cudaMemcpyHostToDevice, c->stream);
        v_cnst0(neq, c->y, c->d_icnstr, c->d_iflag);
        fetch_words(c->d_iflag, c->h_iflag, sizeof(int));
        double bad = (c->h_iflag[0] != INT_MAX) ? 1.0 : 0.0;
        if (c->nranks > 1) {   // every rank must take the same branch
            cudaMemcpyAsync(c->d_scal + 50, &bad, sizeof(double), cudaMemcpyHostToDevice, c->stream);
            allreduce_max(50, 1);
            scal_fetch(50, 1, &bad);
        }
        if (bad != 0.0) { why = "the initial y violates a constraint in ICNSTR"; goto L701; }
    }
    goto L200;

L100:
    if (INFO[1] == 1) goto L200;
    why = "info(1) = -1: the last step ended with an error and no corrective action was taken";
    goto L701;

    // ---- all calls (:1696-1736) ----
L200:
    if (c->neq != neq || !c->arena) { why = "continuation call without a matching initial call"; goto L701; }
    IW[LNSTL] = IW[LNST];
    nzflg = 0;
    rtoli = rtol[0];
    atoli = atol[0];
    for (i64 ii = 1; ii <= neq; ++ii) {
        if (INFO[2] == 1) rtoli = rtol[ii - 1];
        if (INFO[2] == 1) atoli = atol[ii - 1];
        if (rtoli > 0.0 || atoli > 0.0) nzflg = 1;
        if (rtoli < 0.0) { why = "rtol < 0"; goto L701; }
        if (atoli < 0.0) { why = "atol < 0"; goto L701; }
        if (INFO[2] == 0) break;
    }
    if (nzflg == 0) { why = "rtol = atol = 0"; goto L701; }
    // The reference reads RTOL / ATOL in DDAWTS at every step and aliases VT to WT through the work-array pointers
    // (daskr.f:1416-1427, LVT = LWT when info(16) = 0): both must hold on every call, whatever path set the state up
    // (initial call, continuation, dkb_state_import).
    if (INFO[2] == 1) {
        if (!c->d_rtol && (cudaMalloc(&c->d_rtol, sizeof(double) * neq) != cudaSuccess ||
                           cudaMalloc(&c->d_atol, sizeof(double) * neq) != cudaSuccess)) { why = "cannot allocate the tolerance vectors"; goto L701; }
        cudaMemcpyAsync(c->d_rtol, rtol, sizeof(double) * neq, cudaMemcpyHostToDevice, c->stream);
        cudaMemcpyAsync(c->d_atol, atol, sizeof(double) * neq, cudaMemcpyHostToDevice, c->stream);
        cudaStreamSynchronize(c->stream);   // rtol / atol are the caller's pageable arrays
    }
    if (INFO[16] == 0) c->vt = c->wt;
    IW[LLCIWP] = 1;
    if (INFO[1] == 1) goto L400;

    // ---- initial call only (:1744-1936) ----
    tn = *t;
    *idid = 1;
    ier = ewt_set(c, tol, c->y, INFO[16] != 0);
    if (ier != 0) { why = "some element of wt is <= 0"; goto L701; }

    uround = 2.220446049250313e-16;   // D1MACH(4), src/daux.f:2-11
    RW[LROUND] = uround;
    hmin = 4.0 * uround * std::max(fabs(*t), fabs(tout));
    if (INFO[11] != 0) {
        if (INFO[17] == 0)
            RW[LSTOL] = pow(uround, .6667);
        else if (RW[LSTOL] <= 0.0) { why = "stptol <= 0"; goto L701; }
    }
    RW[LEPCON] = 0.33;
    RW[LSQRN] = sqrt((double)c->neq_g);
    RW[LRSQRN] = 1.0 / RW[LSQRN];

    tdist = fabs(tout - *t);
    if (tdist < hmin) { why = "tout too close to t"; goto L701; }

    if (INFO[8] != 0) {
        h0 = RW[LH];
        if ((tout - *t) * h0 < 0.0) { why = "tout behind t"; goto L701; }
        if (h0 == 0.0) { why = "h0 = 0"; goto L701; }
    } else {
        h0 = 0.001 * tdist;
        ypnorm = r_wrms(neq, c->yp, c->vt);
        if (ypnorm > 0.5 / h0) h0 = 0.5 / ypnorm;
        h0 = dsign(h0, tout - *t);
    }
    if (INFO[7] != 0) {
        rh = fabs(h0) / RW[LHMAX];
        if (rh > 1.0) h0 = h0 / rh;
    }
    if (INFO[4] != 0) {
        tstop = RW[LTSTOP];
        if ((tstop - *t) * h0 < 0.0) { why = "tstop behind t"; goto L701; }
        if ((*t + h0 - tstop) * h0 > 0.0) h0 = tstop - *t;
        if ((tstop - tout) * h0 < 0.0) { why = "tstop behind tout"; goto L701; }
    }

    if (INFO[11] != 0) {
        // consistent initial conditions (:1827-1872)
        nwt = 1;
        epconi = RW[LEPIN] * RW[LEPCON];
        tscale = 0.0;
        if (index == 0) tscale = tdist;
        for (;;) {
            k_ddasic(c, cb, tn, INFO[11], &h0, tscale, nwt, idid, &IW[LIWM], RW[LEPLI], RW[LSQRN], RW[LRSQRN],
                     epconi, RW[LSTOL], INFO[15]);
            if (*idid < 0) goto L600;
            if (nwt == 2) break;
            nwt = 2;
            // wt only (vt keeps its first value until :1876, as in the reference)
            ier = ewt_set(c, tol, c->y, false);
            if (ier != 0) { why = "some element of wt is <= 0"; goto L701; }
        }
        if (INFO[14] == 1) {
            *idid = 4;
            h = h0;
            if (INFO[11] == 1) RW[LHOLD] = h0;
            goto L590;
        }
        ier = ewt_set(c, tol, c->y, INFO[16] != 0);
        if (ier != 0) { why = "some element of wt is <= 0"; goto L701; }
        if (INFO[8] != 0) {
            h0 = RW[LH];
        } else {
            h0 = 0.001 * tdist;
            ypnorm = r_wrms(neq, c->yp, c->vt);
            if (ypnorm > 0.5 / h0) h0 = 0.5 / ypnorm;
            h0 = dsign(h0, tout - *t);
        }
        if (INFO[7] != 0) {
            rh = fabs(h0) / RW[LHMAX];
            if (rh > 1.0) h0 = h0 / rh;
        }
        if (INFO[4] != 0) {
            tstop = RW[LTSTOP];
            if ((*t + h0 - tstop) * h0 > 0.0) h0 = tstop - *t;
        }
    }

    h = h0;
    RW[LH] = h;
    v_copy(neq, c->y, c->phi[0]);                 // phi(:,1) = y
    v_scale_to(neq, h, c->yp, c->phi[1]);         // phi(:,2) = h*yprime   (:1917-1920)
    RW[LT0] = *t;
    IW[LIRFND] = 0;
    RW[LPSI] = h;
    RW[LPSI + 1] = 2.0 * h;
    IW[LKOLD] = 1;
    if (nrt != 0) {   // check for a zero of R near the initial T (src/daskr.f:1929-1934)
        k_drchek(c, 1, rtf, nrt, *t, tout, &RW[LPSI - 1], IW[LKOLD], &RW[LR0], &RW[LR0 + nrt], &RW[LR0 + 2 * nrt], jroot,
                 &irt, RW[LROUND], rwork, iwork, rpar, ipar);
        if (irt < 0) { why = "a root function is zero at and near the initial t"; goto L701; }
    }
    goto L500;

    // ---- continuation: stop tests before stepping (:1944-2042) ----
L400:
    uround = RW[LROUND];
    done = false;
    tn = RW[LTN];
    h = RW[LH];
    if (nrt != 0) {   // check for a zero of R near TN (src/daskr.f:1949-1963)
        k_drchek(c, 2, rtf, nrt, tn, tout, &RW[LPSI - 1], IW[LKOLD], &RW[LR0], &RW[LR0 + nrt], &RW[LR0 + 2 * nrt], jroot,
                 &irt, RW[LROUND], rwork, iwork, rpar, ipar);
        if (irt < 0) { why = "root function zero at t and very close to t"; goto L701; }
        if (irt == 1) {
            IW[LIRFND] = 1;
            *idid = 5;
            *t = RW[LT0];
            done = true;
            goto L490;
        }
    }
    if (INFO[7] != 0) {
        rh = fabs(h) / RW[LHMAX];
        if (rh > 1.0) h = h / rh;
    }
    if (*t == tout) { why = "tout == t"; goto L701; }
    if ((*t - tout) * h > 0.0) { why = "tout behind t"; goto L701; }
    if (INFO[4] == 1) goto L430;
    if (INFO[3] == 1) goto L420;
    if ((tn - tout) * h < 0.0) goto L490;
    interp(c, tn, tout, IW[LKOLD], &RW[LPSI - 1]);
    *t = tout;
    *idid = 3;
    done = true;
    goto L490;
L420:
    if ((tn - *t) * h <= 0.0) goto L490;
    if ((tn - tout) * h >= 0.0) {
        interp(c, tn, tout, IW[LKOLD], &RW[LPSI - 1]);
        *t = tout;
        *idid = 3;
    } else {
        interp(c, tn, tn, IW[LKOLD], &RW[LPSI - 1]);
        *t = tn;
        *idid = 1;
    }
    done = true;
    goto L490;
L430:
    tstop = RW[LTSTOP];
    if ((tn - tstop) * h > 0.0) { why = "tstop behind tn"; goto L701; }
    if ((tstop - tout) * h < 0.0) { why = "tstop behind tout"; goto L701; }
    if (INFO[3] == 1) goto L440;
    if ((tn - tout) * h < 0.0) goto L450;
    interp(c, tn, tout, IW[LKOLD], &RW[LPSI - 1]);
    *t = tout;
    *idid = 3;
    done = true;
    goto L490;
L440:
    if ((tn - *t) * h <= 0.0) goto L450;
    if ((tn - tout) * h >= 0.0) {
        interp(c, tn, tout, IW[LKOLD], &RW[LPSI - 1]);
        *t = tout;
        *idid = 3;
    } else {
        interp(c, tn, tn, IW[LKOLD], &RW[LPSI - 1]);
        *t = tn;
        *idid = 1;
    }
    done = true;
    goto L490;
L450:
    if (fabs(tn - tstop) <= 100.0 * uround * (fabs(tn) + fabs(h))) {
        interp(c, tn, tstop, IW[LKOLD], &RW[LPSI - 1]);
        *idid = 2;
        *t = tstop;
        done = true;
        goto L490;
    }
    tnext = tn + h;
    if ((tnext - tstop) * h <= 0.0) goto L490;
    h = tstop - tn;
    RW[LH] = h;
L490:
    if (done) goto L590;

    // ---- step loop (:2053-2168) ----
L500:
    if ((IW[LNST] - IW[LNSTL]) >= c->opt_max_steps) {
        *idid = -1;
        goto L527;
    }
    ier = ewt_set(c, tol, c->phi[0], INFO[16] != 0);
    if (ier != 0) {
        *idid = -3;
        goto L527;
    }
    r = r_wrms(neq, c->phi[0], c->wt) * 100.0 * uround;
    if (r > 1.0) {
        if (INFO[2] == 1) {
            for (i64 ii = 0; ii < neq; ++ii) {
                rtol[ii] = r * rtol[ii];
                atol[ii] = r * atol[ii];
            }
        } else {
            rtol[0] = r * rtol[0];
            atol[0] = r * atol[0];
        }
        *idid = -2;
        goto L527;
    }
    hmin = 4.0 * uround * std::max(fabs(tn), fabs(tout));
    if (INFO[7] != 0) {
        rh = fabs(h) / RW[LHMAX];
        if (rh > 1.0) h = h / rh;
    }

    k_ddstp(c, &tn, cb, &h, &INFO[1], idid, &IW[LIWM], &RW[LALPHA], &RW[LBETA], &RW[LGAMMA], &RW[LPSI], &RW[LSIGMA],
            &RW[LCJ], &RW[LCJOLD], &RW[LHOLD], &RW[LS], hmin, RW[LEPLI], RW[LSQRN], RW[LRSQRN], RW[LEPCON],
            &IW[LPHASE], &IW[LJCALC], INFO[15], &IW[LK], &IW[LKOLD], &IW[LNS], c->s_nonneg, INFO[12]);

L527:
    if (*idid < 0) goto L600;

    // ---- successful step: root check (:2177-2189), then stop tests (:2191-2245) ----
    if (nrt != 0) {
        k_drchek(c, 3, rtf, nrt, tn, tout, &RW[LPSI - 1], IW[LKOLD], &RW[LR0], &RW[LR0 + nrt], &RW[LR0 + 2 * nrt], jroot,
                 &irt, RW[LROUND], rwork, iwork, rpar, ipar);
        if (irt == 1) {
            IW[LIRFND] = 1;
            *idid = 5;
            *t = RW[LT0];
            goto L590;
        }
    }
    if (INFO[4] == 0) {
        if ((tn - tout) * h >= 0.0) {
            interp(c, tn, tout, IW[LKOLD], &RW[LPSI - 1]);
            *t = tout;
            *idid = 3;
            goto L590;
        }
        if (INFO[3] == 0) goto L500;
        *t = tn;
        *idid = 1;
        goto L590;
    }
    tstop = RW[LTSTOP];
    if (fabs(tn - tstop) <= 100.0 * uround * (fabs(tn) + fabs(h))) {
        interp(c, tn, tstop, IW[LKOLD], &RW[LPSI - 1]);
        *t = tstop;
        *idid = 2;
        goto L590;
    }
    if ((tn - tout) * h >= 0.0) {
        interp(c, tn, tout, IW[LKOLD], &RW[LPSI - 1]);
        *t = tout;
        *idid = 3;
        goto L590;
    }
    if (INFO[3] == 0) {
        tnext = tn + h;
        if ((tnext - tstop) * h <= 0.0) goto L500;
        h = tstop - tn;
        goto L500;
    }
    *t = tn;
    *idid = 1;

    // ---- successful returns (:2252-2256) ----
L590:
    RW[LTN] = tn;
    RW[LTLAST] = *t;
    RW[LH] = h;
    copy_out(c, y, yprime, true);
    return;

    // ---- unsuccessful returns other than illegal input (:2263-2386) ----
L600:
    INFO[1] = -1;
    *t = tn;
    RW[LTN] = tn;
    RW[LH] = h;
    if (c->arena) copy_out(c, y, yprime);
    return;

    // ---- illegal input (:2390-2490) ----
L701:
    dkb_fail(DKB_E_ARG, "dkb_daskr: %s", why);
    INFO[1] = -1;
    *idid = -33;
    return;
}
