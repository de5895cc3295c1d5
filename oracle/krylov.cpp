/*
 * oracle/krylov.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restatement of the reference's Krylov linear solver
 * (src/daskr_new.f90:3036-3776: dslvk, dspigm, datv, dorth, dheqr, dhels) and
 * of the norm / weight / interpolation helpers (src/daskr_new.f90:715-868).
 * Statement order and floating-point expression shapes follow the Fortran.
 */
#include "oracle.h"
#include <cmath>
#include <algorithm>

extern "C" {

/* src/daskr_new.f90:828-868 -- two-pass, max-scaled weighted RMS norm */
double o_ddwnrm(long neq, const double *v, const double *rwt)
{
    double vmax = 0.0;
    for (long i = 0; i < neq; ++i)
        if (std::fabs(v[i] * rwt[i]) > vmax) vmax = std::fabs(v[i] * rwt[i]);
    if (vmax <= 0.0) return 0.0;
    double accum = 0.0;
    for (long i = 0; i < neq; ++i) {
        double q = (v[i] * rwt[i]) / vmax;
        accum = accum + q * q;
    }
    return vmax * std::sqrt(accum / (double)neq);
}

/* src/daskr_new.f90:715-749 */
void o_ddawts(long neq, int iwt, const double *rtol, const double *atol,
              const double *y, double *wt)
{
    if (iwt == 0)
        for (long i = 0; i < neq; ++i) wt[i] = rtol[0] * std::fabs(y[i]) + atol[0];
    else
        for (long i = 0; i < neq; ++i) wt[i] = rtol[i] * std::fabs(y[i]) + atol[i];
}

/* src/daskr_new.f90:751-780 */
void o_dinvwt(long neq, double *wt, int *ierr)
{
    for (long i = 0; i < neq; ++i)
        if (wt[i] <= 0.0) {
            *ierr = (int)(i + 1);
            return;
        }
    for (long i = 0; i < neq; ++i) wt[i] = 1.0 / wt[i];
    *ierr = 0;
}

/* src/daskr_new.f90:782-826 */
void o_ddatrp(double t, double tout, double *yout, double *ydotout, long neq,
              int kold, const double *phi, const double *psi)
{
    double dt = tout - t;
    for (long i = 0; i < neq; ++i) yout[i] = phi[i];
    for (long i = 0; i < neq; ++i) ydotout[i] = 0.0;
    double c = 1.0, d = 0.0;
    double gama = dt / psi[0];
    for (int j = 2; j <= kold + 1; ++j) {
        d = d * gama + c / psi[j - 2];
        c = c * gama;
        gama = (dt + psi[j - 2]) / psi[j - 1];
        const double *phij = phi + (long)(j - 1) * neq;
        for (long i = 0; i < neq; ++i) yout[i] = yout[i] + c * phij[i];
        for (long i = 0; i < neq; ++i) ydotout[i] = ydotout[i] + d * phij[i];
    }
}

#define AA(i, j) a[((j) - 1) * lda + ((i) - 1)]

/* Givens coefficients shared by both branches of dheqr
 * (src/daskr_new.f90:3650-3668 and 3698-3715) */
static void givens(double t1, double t2, double *c, double *s)
{
    if (t2 == 0.0) {
        *c = 1.0;
        *s = 0.0;
    } else if (std::fabs(t2) < std::fabs(t1)) {
        double t = t2 / t1;
        *c = 1.0 / std::sqrt(1.0 + t * t);
        *s = -(*c) * t;
    } else {
        double t = t1 / t2;
        *s = -1.0 / std::sqrt(1.0 + t * t);
        *c = -(*s) * t;
    }
}

/* src/daskr_new.f90:3592-3726 */
void o_dheqr(double *a, int lda, int n, double *q, int *info, int ijob)
{
    if (ijob == 1) {
        *info = 0;
        for (int k = 1; k <= n; ++k) {
            int km1 = k - 1, kp1 = k + 1;
            for (int j = 1; j <= km1; ++j) {
                int i = 2 * (j - 1) + 1;
                double t1 = AA(j, k), t2 = AA(j + 1, k);
                double c = q[i - 1], s = q[i];
                AA(j, k) = c * t1 - s * t2;
                AA(j + 1, k) = s * t1 + c * t2;
            }
            int iq = 2 * km1 + 1;
            double t1 = AA(k, k), t2 = AA(kp1, k), c, s;
            givens(t1, t2, &c, &s);
            q[iq - 1] = c;
            q[iq] = s;
            AA(k, k) = c * t1 - s * t2;
            if (AA(k, k) == 0.0) *info = k;
        }
    } else if (ijob >= 2) {
        int nm1 = n - 1;
        for (int k = 1; k <= nm1; ++k) {
            int i = 2 * (k - 1) + 1;
            double t1 = AA(k, n), t2 = AA(k + 1, n);
            double c = q[i - 1], s = q[i];
            AA(k, n) = c * t1 - s * t2;
            AA(k + 1, n) = s * t1 + c * t2;
        }
        *info = 0;
        double t1 = AA(n, n), t2 = AA(n + 1, n), c, s;
        givens(t1, t2, &c, &s);
        int iq = 2 * n - 1;
        q[iq - 1] = c;
        q[iq] = s;
        AA(n, n) = c * t1 - s * t2;
        if (AA(n, n) == 0.0) *info = n;
    }
}

/* src/daskr_new.f90:3728-3776 */
void o_dhels(const double *a, int lda, int n, const double *q, double *b)
{
    for (int k = 1; k <= n; ++k) {
        int kp1 = k + 1, iq = 2 * (k - 1) + 1;
        double c = q[iq - 1], s = q[iq];
        double t1 = b[k - 1], t2 = b[kp1 - 1];
        b[k - 1] = c * t1 - s * t2;
        b[kp1 - 1] = s * t1 + c * t2;
    }
    for (int kb = 1; kb <= n; ++kb) {
        int k = n + 1 - kb;
        b[k - 1] = b[k - 1] / AA(k, k);
        double t = -b[k - 1];
        o_daxpy(k - 1, t, &AA(1, k), b);
    }
}
#undef AA

#define HES(i, j) hes[((j) - 1) * (long)ldhes + ((i) - 1)]
#define V(j) (v + ((long)(j) - 1) * n)

/* src/daskr_new.f90:3522-3590 */
static long g_reorth[2] = {0, 0}; /* test instrumentation: reorthogonalisation branch entered / corrections applied */
void o_reorth_stats(long *out)
{
    out[0] = g_reorth[0];
    out[1] = g_reorth[1];
}

void o_dorth(double *vnew, const double *v, double *hes, long n, int ll,
             int ldhes, int kmp, double *snormw)
{
    double vnorm = o_dnrm2(n, vnew);
    int i0 = std::max(1, ll - kmp + 1);
    for (int i = i0; i <= ll; ++i) {
        HES(i, ll) = o_ddot(n, V(i), vnew);
        double tem = -HES(i, ll);
        o_daxpy(n, tem, V(i), vnew);
    }
    *snormw = o_dnrm2(n, vnew);
    /* absorption test evaluated in FP64 exactly as written (:3575) */
    volatile double probe = vnorm + *snormw / 1000;
    if (probe != vnorm) return;
    g_reorth[0] += 1;

    double sumdsq = 0.0;
    for (int i = i0; i <= ll; ++i) {
        double tem = -o_ddot(n, V(i), vnew);
        volatile double probe2 = HES(i, ll) + tem / 1000;
        if (probe2 == HES(i, ll)) continue;
        HES(i, ll) = HES(i, ll) - tem;
        g_reorth[1] += 1;
        o_daxpy(n, tem, V(i), vnew);
        sumdsq = sumdsq + tem * tem;
    }
    if (sumdsq == 0.0) return;
    double arg = std::max(0.0, (*snormw) * (*snormw) - sumdsq);
    *snormw = std::sqrt(arg);
}
#undef V

/* src/daskr_new.f90:3429-3520 */
void o_datv(long neq, const double *y, double t, const double *ydot,
            const double *savr, const double *v, const double *wght,
            double *ydottemp, o_res_t res, int *ires, o_psol_t psol, double *z,
            double *vtemp, double *rwp, int *iwp, double cj, double epslin,
            int *ierr, int *nres, int *npsol, double *rpar, int *ipar)
{
    int neq32 = (int)neq;
    for (long i = 0; i < neq; ++i) vtemp[i] = v[i] / wght[i];
    for (long i = 0; i < neq; ++i) ydottemp[i] = ydot[i] + vtemp[i] * cj;
    for (long i = 0; i < neq; ++i) z[i] = y[i] + vtemp[i];

    *ires = 0;
    res(&t, z, ydottemp, &cj, vtemp, ires, rpar, ipar);
    *nres = *nres + 1;
    if (*ires < 0) return;

    for (long i = 0; i < neq; ++i) z[i] = vtemp[i] - savr[i];

    *ierr = 0;
    psol(&neq32, &t, y, ydot, savr, ydottemp, &cj, wght, rwp, iwp, z, &epslin, ierr, rpar, ipar);
    *npsol = *npsol + 1;
    if (*ierr != 0) return;

    for (long i = 0; i < neq; ++i) z[i] = z[i] * wght[i];
}

#define VV(j) (v + ((long)(j) - 1) * neq)

/* src/daskr_new.f90:3167-3427 */
void o_dspigm(long neq, double t, const double *y, const double *ydot,
              const double *savr, double *r, const double *wght, int maxl,
              int kmp, double epslin, double cj, o_res_t res, int *ires,
              int *nres, o_psol_t psol, int *npsol, double *z, double *v,
              double *hes, double *q, int *lgmr, double *rwp, int *iwp,
              double *wk, double *dl, double *rhok, int *iflag, int irst,
              int nrsts, double *rpar, int *ipar)
{
    int neq32 = (int)neq;
    int ierr = 0, info = 0, ll = 0, llp1, maxlp1 = maxl + 1;
    const int ldhes = maxl + 1;
    double c, dlnorm, prod, rho = 0.0, rnorm, s, snormw = 0.0, tem;

    *iflag = 0;
    *lgmr = 0;
    *npsol = 0;
    *nres = 0;

    for (long i = 0; i < neq; ++i) z[i] = 0.0;

    if (nrsts == 0) {
        psol(&neq32, &t, y, ydot, savr, wk, &cj, wght, rwp, iwp, r, &epslin, &ierr, rpar, ipar);
        *npsol = 1;
        if (ierr != 0) goto L300;
        for (long i = 0; i < neq; ++i) VV(1)[i] = r[i] * wght[i];
    } else {
        for (long i = 0; i < neq; ++i) VV(1)[i] = r[i];
    }

    rnorm = o_dnrm2(neq, v);
    if (rnorm <= epslin) {
        *rhok = rnorm;
        return;
    }
    tem = 1.0 / rnorm;
    o_dscal(neq, tem, VV(1));

    for (int i = 0; i < (maxl + 1) * maxl; ++i) hes[i] = 0.0;

    prod = 1.0;
    for (ll = 1; ll <= maxl; ++ll) {
        *lgmr = ll;
        o_datv(neq, y, t, ydot, savr, VV(ll), wght, z, res, ires, psol, VV(ll + 1), wk,
               rwp, iwp, cj, epslin, &ierr, nres, npsol, rpar, ipar);
        if (*ires < 0) return;
        if (ierr != 0) goto L300;

        o_dorth(VV(ll + 1), v, hes, neq, ll, maxlp1, kmp, &snormw);
        HES(ll + 1, ll) = snormw;

        o_dheqr(hes, maxlp1, ll, q, &info, ll);
        if (info == ll) goto L120;

        prod = prod * q[2 * ll - 1];
        rho = std::fabs(prod * rnorm);
        if ((ll > kmp) && (kmp < maxl)) {
            if (ll == kmp + 1) {
                o_dcopy(neq, VV(1), dl);
                for (int i = 1; i <= kmp; ++i) {
                    int ip1 = i + 1, i2 = i * 2;
                    s = q[i2 - 1];
                    c = q[i2 - 2];
                    for (long k = 0; k < neq; ++k) dl[k] = s * dl[k] + c * VV(ip1)[k];
                }
            }
            s = q[2 * ll - 1];
            c = q[2 * ll - 2] / snormw;
            llp1 = ll + 1;
            for (long k = 0; k < neq; ++k) dl[k] = s * dl[k] + c * VV(llp1)[k];
            dlnorm = o_dnrm2(neq, dl);
            rho = rho * dlnorm;
        }

        if (rho <= epslin) goto L200;
        if (ll == maxl) goto L100;

        tem = 1.0 / snormw;
        o_dscal(neq, tem, VV(ll + 1));
    }

L100:
    if (rho < rnorm) goto L150;

L120:
    *iflag = 2;
    for (long i = 0; i < neq; ++i) z[i] = 0.0;
    return;

L150:
    *iflag = 1;
    if (irst > 0) {
        if (kmp == maxl) {
            o_dcopy(neq, VV(1), dl);
            int maxlm1 = maxl - 1;
            for (int i = 1; i <= maxlm1; ++i) {
                int ip1 = i + 1, i2 = i * 2;
                s = q[i2 - 1];
                c = q[i2 - 2];
                for (long k = 0; k < neq; ++k) dl[k] = s * dl[k] + c * VV(ip1)[k];
            }
            s = q[2 * maxl - 1];
            c = q[2 * maxl - 2] / snormw;
            for (long k = 0; k < neq; ++k) dl[k] = s * dl[k] + c * VV(maxlp1)[k];
        }
        tem = rnorm * prod;
        o_dscal(neq, tem, dl);
    }

L200:
    ll = *lgmr;
    llp1 = ll + 1;
    for (int i = 0; i < llp1; ++i) r[i] = 0.0;
    r[0] = rnorm;
    o_dhels(hes, maxlp1, ll, q, r);
    for (long i = 0; i < neq; ++i) z[i] = 0.0;
    for (int i = 1; i <= ll; ++i) o_daxpy(neq, r[i - 1], VV(i), z);
    for (long i = 0; i < neq; ++i) z[i] = z[i] / wght[i];
    *rhok = rho;
    return;

L300:
    if (ierr < 0) *iflag = -1;
    if (ierr > 0) *iflag = 3;
}
#undef VV
#undef HES

/* iwork slots, src/daskr_new.f90:86-145 (1-based) */
enum { IW_NRE = 12, IW_NCFL = 16, IW_NLI = 20, IW_NPS = 21, IW_MITER = 23, IW_MAXL = 24,
       IW_KMP = 25, IW_NRMAX = 26, IW_LRWP = 29, IW_LIWP = 30 };

/* src/daskr_new.f90:3036-3165.  rwm/iwm are the WM / IWORK bases; the
 * work-vector carving is the reference's (:3112-3119) with 64-bit offsets.
 * iwm(29) holds LOCWP as a 32-bit value in the reference; oracle callers that
 * need more than 2^31 words are out of its range (CPU sizes only). */
void o_dslvk(long neq, const double *y, double t, const double *ydot,
             const double *savr, double *x, double *ewt, double *rwm, int *iwm,
             o_res_t res, int *ires, o_psol_t psol, int *iersl, double cj,
             double epslin, double sqrtn, double rsqrtn, double *rhok,
             double *rpar, int *ipar)
{
    const int irst = 1; /* :3093, never changed */
    int *IWM = iwm - 1;
    double *RWM = rwm - 1;
    int liwp = IWM[IW_LIWP], nli = IWM[IW_NLI], nps = IWM[IW_NPS], ncfl = IWM[IW_NCFL],
        nre = IWM[IW_NRE], lrwp = IWM[IW_LRWP], maxl = IWM[IW_MAXL], kmp = IWM[IW_KMP],
        nrmax = IWM[IW_NRMAX];
    int iflag, nrsts, lgmr = 0, npsol = 0, nres = 0;
    *iersl = 0;
    *ires = 0;

    int maxlp1 = maxl + 1;
    long lv = 1;
    long lr = lv + neq * maxl;
    long lhes = lr + neq + 1;
    long lq = lhes + (long)maxl * maxlp1;
    long lwk = lq + 2 * maxl;
    long ldl = lwk + std::min(1, maxl - kmp) * neq;
    long lz = ldl + neq;
    o_dscal(neq, rsqrtn, ewt);
    o_dcopy(neq, x, &RWM[lr]);
    for (long i = 0; i < neq; ++i) x[i] = 0.0;

    iflag = 1;
    nrsts = -1;
    while ((iflag == 1) && (nrsts < nrmax) && (*ires == 0)) {
        nrsts = nrsts + 1;
        if (nrsts > 0) o_dcopy(neq, &RWM[ldl], &RWM[lr]);
        o_dspigm(neq, t, y, ydot, savr, &RWM[lr], ewt, maxl, kmp, epslin, cj, res, ires,
                 &nres, psol, &npsol, &RWM[lz], &RWM[lv], &RWM[lhes], &RWM[lq], &lgmr,
                 &RWM[lrwp], &IWM[liwp], &RWM[lwk], &RWM[ldl], rhok, &iflag, irst, nrsts,
                 rpar, ipar);
        nli = nli + lgmr;
        nps = nps + npsol;
        nre = nre + nres;
        for (long i = 0; i < neq; ++i) x[i] = x[i] + RWM[lz + i];
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
    o_dscal(neq, sqrtn, ewt);
}

} /* extern "C" */
