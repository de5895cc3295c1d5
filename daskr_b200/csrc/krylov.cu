// krylov.cu -- device-resident restarted GMRES: dslvk / dspigm / datv / dorth
// (src/daskr_new.f90:3036-3590) with dheqr / dhels (:3592-3776) on host scalars.
//
// Control flow, tolerances, flags and counters are the reference's; every
// length-NEQ statement is a kernel on vectors that never leave HBM.  Per Krylov
// iteration the host sees one 8*(ll+2)-byte readback (the new Hessenberg column
// and the two norms) -- the Gram-Schmidt coefficients themselves stay in device
// scalar slots between the fused axpy+dot passes, so there is no host
// round-trip inside dorth.
//
// Two forms of the Jacobian-vector product:
//  * fused (built-in problems, reaction preconditioner only): one kernel does
//    z = wght * P^-1 ( G(y + v/wght, ydot + cj v/wght) - savr )  (K3);
//  * generic (user callbacks with device pointers): the five statements of
//    datv:3493-3518 as separate kernels around res() and psol().
#include "dkb_internal.h"
#include <math.h>
#include <algorithm>

// scalar slot map in d_scal
enum { S_VNORM = 0, S_H = 1 /* .. S_H+maxl-1 */, S_SNORM = 1 + DKB_MAXL_MAX, S_RNORM = 2 + DKB_MAXL_MAX,
       S_TMP = 3 + DKB_MAXL_MAX };

// iwork slots (src/daskr_new.f90:86-145, 1-based)
enum { IW_NRE = 12, IW_NCFL = 16, IW_NLI = 20, IW_NPS = 21, IW_MAXL = 24, IW_KMP = 25, IW_NRMAX = 26 };

#define HES(i, j) hes[((j) - 1) * ldh + ((i) - 1)]

static void givens(double t1, double t2, double *c, double *s)
{
    // src/daskr_new.f90:3698-3715
    if (t2 == 0.0) {
        *c = 1.0;
        *s = 0.0;
    } else if (fabs(t2) < fabs(t1)) {
        double t = t2 / t1;
        *c = 1.0 / sqrt(1.0 + t * t);
        *s = -(*c) * t;
    } else {
        double t = t1 / t2;
        *s = -1.0 / sqrt(1.0 + t * t);
        *c = -(*s) * t;
    }
}

// dheqr (:3592-3726) as dspigm calls it: column n has just been appended (ijob = n;
// n = 1 is the "fresh factorisation" branch, which for one column is the same update)
static int h_dheqr_update(double *hes, int ldh, int n, double *q)
{
    for (int k = 1; k <= n - 1; ++k) {
        int i = 2 * (k - 1);
        double t1 = HES(k, n), t2 = HES(k + 1, n);
        double c = q[i], s = q[i + 1];
        HES(k, n) = c * t1 - s * t2;
        HES(k + 1, n) = s * t1 + c * t2;
    }
    double t1 = HES(n, n), t2 = HES(n + 1, n), c, s;
    givens(t1, t2, &c, &s);
    q[2 * n - 2] = c;
    q[2 * n - 1] = s;
    HES(n, n) = c * t1 - s * t2;
    return (HES(n, n) == 0.0) ? n : 0;
}

// dhels (:3728-3776)
static void h_dhels(const double *hes, int ldh, int n, const double *q, double *b)
{
    for (int k = 1; k <= n; ++k) {
        double c = q[2 * k - 2], s = q[2 * k - 1];
        double t1 = b[k - 1], t2 = b[k];
        b[k - 1] = c * t1 - s * t2;
        b[k] = s * t1 + c * t2;
    }
    for (int kb = 1; kb <= n; ++kb) {
        int k = n + 1 - kb;
        b[k - 1] = b[k - 1] / HES(k, k);
        double t = -b[k - 1];
        if (t != 0.0)
            for (int i = 1; i <= k - 1; ++i) b[i - 1] = b[i - 1] + t * HES(i, k);
    }
}

bool builtin_fused(dkb_res_t res, dkb_psol_t psol, const int *ipar)
{
    DkbCtx *c = g_ctx;
    if (!c->opt_fused || !c->rbd_set) return false;
    if (c->problem == DKB_PROB_WEB && res == dkb_web_res && psol == dkb_web_psol)
        return ipar != nullptr && (ipar[0] == 1 || ipar[0] == 3) && ipar[1] == 0;   // jpre = 1 or 3, no block grouping
    if (c->problem == DKB_PROB_HEAT && res == dkb_heat_res && psol == dkb_heat_psol) return true;
    return false;
}

struct GmresArgs {
    i64 neq;
    double t, cj, epslin, rsqrtn;
    const double *y, *ydot, *savr, *wt;
    double *x;
    dkb_res_t res;
    dkb_psol_t psol;
    double *rpar;
    int *ipar;
    bool fused;
    int jpre;   // food web: ipar(1)
};

// datv, src/daskr_new.f90:3429-3520: vnext = D^-1 P^-1 (dF/dy) D v.  The basis vector arrives unnormalised with its
// factor vs = 1/||.||: every kernel forms vs*v[i] on load -- the product the reference's dscal stored (dspigm:3295, 3358)
// -- which saves the separate normalisation pass over v(:,ll+1) (16 bytes per unknown and Krylov iteration).
static void k_datv(const GmresArgs &A, const double *v, double vs, double *vnext, int *ires, int *ierr, int *nres, int *npsol)
{
    DkbCtx *c = g_ctx;
    ProfScope ps(PROF_DATV);
    *ires = 0;
    *ierr = 0;
    if (A.fused) {
        if (c->nranks > 1) halo_exchange(const_cast<double *>(v));
        if (c->problem == DKB_PROB_WEB) {
            if (web_datv_fused(v, vs, A.wt, A.rsqrtn, A.y, A.ydot, A.savr, A.cj, c->wp, c->iwp, vnext, A.jpre) != 0) *ierr = -1;
        } else
            heat_datv_fused(v, vs, A.wt, A.rsqrtn, A.y, A.ydot, A.savr, A.cj, c->wp, vnext);
        *nres += 1;
        *npsol += 1;
        return;
    }
    const int neq32 = (int)A.neq;
    double *vtemp = c->wk, *ydottemp = c->z;
    v_datv_pre(A.neq, v, vs, A.wt, A.rsqrtn, A.y, A.ydot, A.cj, vtemp, ydottemp, vnext);
    A.res(&A.t, vnext, ydottemp, &A.cj, vtemp, ires, A.rpar, A.ipar);
    *nres += 1;
    if (*ires < 0) return;
    v_sub(A.neq, vtemp, A.savr, vnext);
    A.psol(&neq32, &A.t, A.y, A.ydot, A.savr, ydottemp, &A.cj, c->wght, c->wp, c->iwp, vnext, &A.epslin, ierr,
           A.rpar, A.ipar);
    *npsol += 1;
    if (*ierr != 0) return;
    v_mul(A.neq, vnext, c->wght, vnext);
}

// dorth, src/daskr_new.f90:3522-3590.  Fills column ll of hes, returns snormw.
static double k_dorth(i64 n, double *vnew, double *const *v, const double *vs, double *hes, int ldh, int ll, int kmp)
{
    ProfScope ps(PROF_MGS);
    const int i0 = std::max(1, ll - kmp + 1);
    // vnorm^2 and the first coefficient in one pass, then fused (axpy_i , dot_{i+1}) passes,
    // the last one fused with the norm of the result.
    r_dot2(n, vnew, v[i0 - 1], vs[i0 - 1], S_VNORM);   // S_VNORM, S_VNORM+1 (= S_H + 0 scratch)
    // r_dot2 wrote [vnorm^2, h_i0] to slots S_VNORM, S_VNORM+1; S_H == S_VNORM+1 holds h_i0.
    int hslot = S_H;
    for (int i = i0; i <= ll - 1; ++i) {
        r_axpy_dot(n, vnew, v[i - 1], vs[i - 1], hslot, v[i], vs[i], hslot + 1);
        hslot += 1;
    }
    r_axpy_nrm(n, vnew, v[ll - 1], vs[ll - 1], hslot, S_SNORM);
    double buf[DKB_MAXL_MAX + 2];
    scal_fetch(S_VNORM, S_SNORM - S_VNORM + 1, buf);   // one 144-byte readback per Krylov iteration
    const double sn2 = buf[S_SNORM - S_VNORM];
    const double vnorm = sqrt(buf[0]);
    for (int i = i0; i <= ll; ++i) HES(i, ll) = buf[1 + (i - i0)];
    double snormw = sqrt(sn2);

    // conditional reorthogonalisation (:3575-3588), evaluated in FP64 exactly as written
    volatile double probe = vnorm + snormw / 1000;
    if (probe != vnorm) return snormw;
    g_ctx->reorth_entered += 1;
    double sumdsq = 0.0;
    for (int i = i0; i <= ll; ++i) {
        double d;
        r_dot_s(n, v[i - 1], vs[i - 1], vnew, S_TMP);
        scal_fetch(S_TMP, 1, &d);
        double tem = -d;
        volatile double probe2 = HES(i, ll) + tem / 1000;
        if (probe2 == HES(i, ll)) continue;
        HES(i, ll) = HES(i, ll) - tem;
        g_ctx->reorth_corrections += 1;
        v_axpy_s(n, tem, v[i - 1], vs[i - 1], vnew);
        sumdsq = sumdsq + tem * tem;
    }
    if (sumdsq == 0.0) return snormw;
    double arg = std::max(0.0, snormw * snormw - sumdsq);
    return sqrt(arg);
}

// dspigm, src/daskr_new.f90:3167-3427.  On return *wrote_x says whether x was updated
// (x = x + z with z the new correction; first = initial pass of dslvk where x == 0).
static void k_dspigm(const GmresArgs &A, int maxl, int kmp, int *ires, int *nres, int *npsol, double *hes,
                     double *q, int *lgmr, double *rhok, int *iflag, int nrsts)
{
    DkbCtx *c = g_ctx;
    const i64 neq = A.neq;
    const int ldh = maxl + 1;
    const int neq32 = (int)neq;
    int ierr = 0, info = 0, ll = 0;
    double prod, rho = 0.0, rnorm, snormw = 0.0, sumsq, dlnorm;
    double **v = c->v;
    double vs[DKB_MAXL_MAX + 1];   // vs[k] = 1/||v_k||: the basis is stored unnormalised, consumers multiply on load

    *iflag = 0;
    *lgmr = 0;
    *npsol = 0;
    *nres = 0;
    const int first = (nrsts == 0);

    // v1 = wght * P^-1 r  (nrsts == 0) or the restart residual dl (already scaled)
    if (nrsts == 0) {
        ProfScope ps(PROF_PSOL0);
        if (A.fused) {
            if (c->ns == 1) {
                lu_psol_scaled(c->wp, c->iwp, 1, c->npts, A.x, A.wt, A.rsqrtn, v[0], S_RNORM);
            } else if (A.jpre == 3) {   // A_S (Gauss-Seidel, in place on a copy of the rhs) then A_R
                double *r = v[maxl], *gsout = r;
                v_copy(neq, A.x, r);
                if (web_gauss_seidel_to(1.0 / A.cj, r, c->wk, &gsout) != 0) ierr = -1;
                tpp_psol(c->wp, c->iwp, c->ns, c->npts, gsout, nullptr, A.wt, A.rsqrtn, v[0], S_RNORM);
            } else {
                tpp_psol(c->wp, c->iwp, c->ns, c->npts, A.x, nullptr, A.wt, A.rsqrtn, v[0], S_RNORM);
            }
        } else {
            double *r = v[maxl];   // r is v(:,maxl+1), dslvk:3114
            v_copy(neq, A.x, r);
            A.psol(&neq32, &A.t, A.y, A.ydot, A.savr, c->wk, &A.cj, c->wght, c->wp, c->iwp, r, &A.epslin, &ierr,
                   A.rpar, A.ipar);
            if (ierr == 0) {
                v_mul(neq, r, c->wght, v[0]);
                r_dot(neq, v[0], v[0], S_RNORM);
            }
        }
        *npsol = 1;
        if (ierr != 0) goto L300;
    } else {
        ProfScope ps(PROF_VEC);
        v_copy(neq, c->dl, v[0]);
        r_dot(neq, v[0], v[0], S_RNORM);
    }
    scal_fetch(S_RNORM, 1, &sumsq);
    rnorm = sqrt(sumsq);
    if (rnorm <= A.epslin) {   // :3290-3293, z = 0
        if (first) v_zero(neq, A.x);
        *rhok = rnorm;
        return;
    }
    vs[0] = 1.0 / rnorm;   // :3294-3295 (tem = one/rnorm; dscal)
    for (int i = 0; i < ldh * maxl; ++i) hes[i] = 0.0;

    prod = 1.0;
    for (ll = 1; ll <= maxl; ++ll) {
        *lgmr = ll;
        ProfScope kit(PROF_KIT);   // one Arnoldi iteration: datv + dorth (+ the normalisation of v(:,ll+1))
        c->ll_hist[ll] += 1;
        k_datv(A, v[ll - 1], vs[ll - 1], v[ll], ires, &ierr, nres, npsol);
        if (*ires < 0) {
            if (first) v_zero(neq, A.x);
            return;
        }
        if (ierr != 0) goto L300;

        snormw = k_dorth(neq, v[ll], v, vs, hes, ldh, ll, kmp);
        HES(ll + 1, ll) = snormw;

        info = h_dheqr_update(hes, ldh, ll, q);
        if (info == ll) goto L120;

        prod = prod * q[2 * ll - 1];
        rho = fabs(prod * rnorm);
        if ((ll > kmp) && (kmp < maxl)) {
            // incomplete orthogonalisation: update dl and use its norm (:3329-3350)
            if (ll == kmp + 1) {
                v_scale_to(neq, vs[0], v[0], c->dl);
                for (int i = 1; i <= kmp; ++i) v_lin2_s(neq, q[2 * i - 1], c->dl, q[2 * i - 2], v[i], vs[i], c->dl);
            }
            v_lin2(neq, q[2 * ll - 1], c->dl, q[2 * ll - 2] / snormw, v[ll], c->dl);
            dlnorm = r_nrm2(neq, c->dl);
            rho = rho * dlnorm;
        }
        if (rho <= A.epslin) goto L200;
        if (ll == maxl) goto L100;
        vs[ll] = 1.0 / snormw;   // :3357-3359 (tem = one/snormw; dscal): applied by the consumers of v(:,ll+1)
    }

L100:
    if (rho < rnorm) goto L150;
L120:
    *iflag = 2;
    if (first) v_zero(neq, A.x);   // z = 0
    return;

L150:
    *iflag = 1;
    // restart residual dl (irst > 0 always, dslvk:3093)
    {
        ProfScope ps(PROF_VEC);
        if (kmp == maxl)
            v_gmres_dl(neq, maxl, v, vs, q, snormw, rnorm * prod, c->dl);
        else
            v_scale(neq, rnorm * prod, c->dl);
    }

L200: {
    ll = *lgmr;
    double r[DKB_MAXL_MAX + 2];
    for (int i = 0; i <= ll; ++i) r[i] = 0.0;
    r[0] = rnorm;
    h_dhels(hes, ldh, ll, q, r);
    ProfScope ps(PROF_VEC);
    v_gmres_solution(neq, ll, v, vs, r, A.wt, A.rsqrtn, A.x, first);
    *rhok = rho;
    return;
}

L300:
    if (first) v_zero(neq, A.x);
    if (ierr < 0) *iflag = -1;
    if (ierr > 0) *iflag = 3;
}

// dslvk, src/daskr_new.f90:3036-3165
void k_dslvk(i64 neq, const double *y, double t, const double *ydot, const double *savr, double *x,
             const double *ewt, int *iwm, dkb_res_t res, int *ires, dkb_psol_t psol, int *iersl, double cj,
             double epslin, double sqrtn, double rsqrtn, double *rhok, double *rpar, int *ipar)
{
    DkbCtx *c = g_ctx;
    (void)sqrtn;
    int *IWM = iwm - 1;
    int nli = IWM[IW_NLI], nps = IWM[IW_NPS], ncfl = IWM[IW_NCFL], nre = IWM[IW_NRE];
    const int maxl = IWM[IW_MAXL], kmp = IWM[IW_KMP], nrmax = IWM[IW_NRMAX];
    double hes[(DKB_MAXL_MAX + 1) * DKB_MAXL_MAX], q[2 * DKB_MAXL_MAX];
    int iflag = 1, nrsts = -1, lgmr = 0, npsol = 0, nres = 0;
    *iersl = 0;
    *ires = 0;

    GmresArgs A;
    A.neq = neq;
    A.t = t;
    A.cj = cj;
    A.epslin = epslin;
    A.rsqrtn = rsqrtn;
    A.y = y;
    A.ydot = ydot;
    A.savr = savr;
    A.wt = ewt;
    A.x = x;
    A.res = res;
    A.psol = psol;
    A.rpar = rpar;
    A.ipar = ipar;
    A.fused = builtin_fused(res, psol, ipar);
    A.jpre = (A.fused && c->problem == DKB_PROB_WEB) ? ipar[0] : 1;

    // The reference scales ewt by 1/sqrt(N) in place (:3120) and back on exit (:3163).  Here the
    // weights stay untouched: fused kernels form wght = rsqrtn*wt on the fly (same product,
    // bit for bit); only user callbacks get a materialised copy.
    if (!A.fused) v_scale_to(neq, rsqrtn, ewt, c->wght);
    if (A.fused && c->nranks > 1) {
        // ghost rows of y and wt are constant during this solve; v's are exchanged per datv
        halo_exchange(const_cast<double *>(y));
        halo_exchange(const_cast<double *>(ewt));
    }

    while ((iflag == 1) && (nrsts < nrmax) && (*ires == 0)) {
        nrsts = nrsts + 1;
        k_dspigm(A, maxl, kmp, ires, &nres, &npsol, hes, q, &lgmr, rhok, &iflag, nrsts);
        nli = nli + lgmr;
        nps = nps + npsol;
        nre = nre + nres;
    }

    if (*ires < 0) {
        ncfl = ncfl + 1;
    } else if (iflag != 0) {
        ncfl = ncfl + 1;
        if (iflag > 0) *iersl = 1;
        if (iflag < 0) *iersl = -1;
    }
    IWM[IW_NLI] = nli;
    IWM[IW_NPS] = nps;
    IWM[IW_NCFL] = ncfl;
    IWM[IW_NRE] = nre;
}
