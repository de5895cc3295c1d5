/*
 * oracle/daskr_driver.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restatement of the host logic that drives the Krylov hot path:
 *   DASKR  (src/daskr.f:1402-2493, Krylov branches info(12)=1 only),
 *   DDASIC / DYYPNW (src/daskr.f:2911-3131),
 *   ddstp  (src/daskr_new.f90:149-567),  dcnstr/dcnst0 (:569-713),
 *   ddasik / dnsik / dlinsk / dfnrmk (:2004-2658),
 *   dnedk / dnsk (:2660-3034).
 * Not restated (out of the hot-path scope): the
 * direct dense/banded solvers and the XERRWD
 * message printer (idid codes are returned, messages are not printed).
 * The work arrays keep the reference's layout so that
 * rwork(1:60) / iwork(1:40) can be compared slot for slot.
 */
#include "oracle.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <vector>
#include <algorithm>

/* iwork pointers, src/daskr.f:1414-1420 */
enum { LML = 1, LMU = 2, LMTYPE = 4, LIWM = 1, LMXORD = 3, LJCALC = 5, LPHASE = 6, LK = 7,
       LKOLD = 8, LNS = 9, LNSTL = 10, LNST = 11, LNRE = 12, LNJE = 13, LETF = 14,
       LNCFN = 15, LNCFL = 16, LNIW = 17, LNRW = 18, LNNI = 19, LNLI = 20, LNPS = 21,
       LNPD = 22, LMITER = 23, LMAXL = 24, LKMP = 25, LNRMAX = 26, LLNWP = 27, LLNIWP = 28,
       LLOCWP = 29, LLCIWP = 30, LKPRIN = 31, LMXNIT = 32, LMXNJ = 33, LMXNH = 34,
       LLSOFF = 35, LNRTE = 36, LIRFND = 37, LICNS = 41 };
/* rwork pointers, src/daskr.f:1424-1427 */
enum { LTSTOP = 1, LHMAX = 2, LH = 3, LTN = 4, LCJ = 5, LCJOLD = 6, LHOLD = 7, LS = 8,
       LROUND = 9, LEPLI = 10, LSQRN = 11, LRSQRN = 12, LEPCON = 13, LSTOL = 14, LEPIN = 15,
       LALPHA = 21, LBETA = 27, LGAMMA = 33, LPSI = 39, LSIGMA = 45, LT0 = 51, LTLAST = 52,
       LDELTA = 61 };

static inline double dsign(double a, double b) { return b >= 0.0 ? std::fabs(a) : -std::fabs(a); }

/* src/daskr_new.f90:663-713 */
static void dcnst0(long neq, const double *y, const int *icnstr, int *iret)
{
    *iret = 0;
    for (long i = 0; i < neq; ++i) {
        if (icnstr[i] == 2) { if (y[i] <= 0.0) { *iret = (int)(i + 1); break; } }
        else if (icnstr[i] == 1) { if (y[i] < 0.0) { *iret = (int)(i + 1); break; } }
        else if (icnstr[i] == -1) { if (y[i] > 0.0) { *iret = (int)(i + 1); break; } }
        else if (icnstr[i] == -2) { if (y[i] >= 0.0) { *iret = (int)(i + 1); break; } }
    }
}

/* src/daskr_new.f90:569-661 */
static void dcnstr(long neq, const double *y, const double *ynew, const int *icnstr,
                   double *tau, double rlx, int *iret, int *ivar)
{
    const double fac = 0.6, fac2 = 0.9;
    double rdy, rdymx = 0.0;
    *iret = 0;
    *ivar = 0;
    for (long i = 0; i < neq; ++i) {
        if (icnstr[i] == 2) {
            rdy = std::fabs((ynew[i] - y[i]) / y[i]);
            if (rdy > rdymx) { rdymx = rdy; *ivar = (int)(i + 1); }
            if (ynew[i] <= 0.0) { *tau = fac * (*tau); *ivar = (int)(i + 1); *iret = 1; return; }
        } else if (icnstr[i] == 1) {
            if (ynew[i] < 0.0) { *tau = fac * (*tau); *ivar = (int)(i + 1); *iret = 1; return; }
        } else if (icnstr[i] == -1) {
            if (ynew[i] > 0.0) { *tau = fac * (*tau); *ivar = (int)(i + 1); *iret = 1; return; }
        } else if (icnstr[i] == -2) {
            rdy = std::fabs((ynew[i] - y[i]) / y[i]);
            if (rdy > rdymx) { rdymx = rdy; *ivar = (int)(i + 1); }
            if (ynew[i] >= 0.0) { *tau = fac * (*tau); *ivar = (int)(i + 1); *iret = 1; return; }
        }
    }
    if (rdymx >= rlx) {
        *tau = fac2 * (*tau) * rlx / rdymx;
        *iret = 1;
    }
}

/* src/daskr.f:3078-3131 */
static void dyypnw(long neq, const double *y, const double *yprime, double cj, double rl,
                   const double *p, int icopt, const int *id, double *ynew, double *ypnew)
{
    if (icopt == 1) {
        for (long i = 0; i < neq; ++i) {
            if (id[i] < 0) {
                ynew[i] = y[i] - rl * p[i];
                ypnew[i] = yprime[i];
            } else {
                ynew[i] = y[i];
                ypnew[i] = yprime[i] - rl * cj * p[i];
            }
        }
    } else {
        for (long i = 0; i < neq; ++i) {
            ynew[i] = y[i] - rl * p[i];
            ypnew[i] = yprime[i];
        }
    }
}

/* src/daskr_new.f90:2568-2658 */
static void dfnrmk(long neq, const double *y, double t, const double *ydot, double *savr,
                   double *r, double cj, double tscale, double *wght, double sqrtn,
                   double rsqrtn, o_res_t res, int *ires, o_psol_t psol, int irin, int *ierr,
                   double *rnorm, double epslin, double *rwp, int *iwp, double *wk,
                   double *rpar, int *ipar)
{
    int neq32 = (int)neq;
    if (irin == 0) {
        *ires = 0;
        res(&t, y, ydot, &cj, savr, ires, rpar, ipar);
        if (*ires < 0) return;
    }
    o_dcopy(neq, savr, r);
    o_dscal(neq, rsqrtn, wght);
    *ierr = 0;
    psol(&neq32, &t, y, ydot, savr, wk, &cj, wght, rwp, iwp, r, &epslin, ierr, rpar, ipar);
    o_dscal(neq, sqrtn, wght);
    if (*ierr != 0) return;
    *rnorm = o_ddwnrm(neq, r, wght);
    if (tscale > 0.0) *rnorm = (*rnorm) * tscale * std::fabs(cj);
}

/* src/daskr_new.f90:2365-2566 (kprint messages omitted) */
static void dlinsk(long neq, double *y, double t, double *ydot, double *savr, double cj,
                   double tscale, double *p, double *pnorm, double *wght, double sqrtn,
                   double rsqrtn, int lsoff, double stptol, int *iret, o_res_t res,
                   int *ires, o_psol_t psol, double *rwm, int *iwm, double rhok,
                   double *fnorm, int icopt, const int *idy, double *rwp, int *iwp,
                   double *r, double epslin, double *ynew, double *ydotnew, double *pwk,
                   int icnflg, int *icnstr, double rlx, double *rpar, int *ipar)
{
    (void)rwm; (void)rhok;
    const double alpha = 1e-4;
    int *IWM = iwm - 1;
    int ierr = 0, ivar;
    double f1norm, f1normp, fnormp = 0.0, ratio, ratio1, rl, rlmin, slpi, tau;

    f1norm = ((*fnorm) * (*fnorm)) / 2;
    ratio = 1.0;
    tau = *pnorm;
    rl = 1.0;

    if (icnflg != 0) {
        for (;;) {
            dyypnw(neq, y, ydot, cj, rl, p, icopt, idy, ynew, ydotnew);
            dcnstr(neq, y, ynew, icnstr, &tau, rlx, iret, &ivar);
            if (*iret != 1) break;
            ratio1 = tau / (*pnorm);
            ratio = ratio * ratio1;
            for (long i = 0; i < neq; ++i) p[i] = p[i] * ratio1;
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
        dyypnw(neq, y, ydot, cj, rl, p, icopt, idy, ynew, ydotnew);
        dfnrmk(neq, ynew, t, ydotnew, savr, r, cj, tscale, wght, sqrtn, rsqrtn, res, ires,
               psol, 0, &ierr, &fnormp, epslin, rwp, iwp, pwk, rpar, ipar);
        IWM[LNRE] = IWM[LNRE] + 1;
        if (*ires >= 0) IWM[LNPS] = IWM[LNPS] + 1;
        if (*ires != 0 || ierr != 0) {
            *iret = 2;
            return;
        }
        if (lsoff != 1) {
            f1normp = (fnormp * fnormp) / 2;
            if (f1normp > f1norm + alpha * slpi * rl) {
                /* alpha-condition not satisfied: backtrack */
                if (rl < rlmin) {
                    *iret = 1;
                    return;
                }
                rl = rl / 2;
                continue;
            }
        }
        *iret = 0;
        for (long i = 0; i < neq; ++i) y[i] = ynew[i];
        for (long i = 0; i < neq; ++i) ydot[i] = ydotnew[i];
        *fnorm = fnormp;
        return;
    }
}

/* src/daskr_new.f90:2175-2363 */
static void dnsik(double t, double *y, double *ydot, long neq, int icopt, const int *idy,
                  o_res_t res, o_psol_t psol, double *wt, double *rpar, int *ipar,
                  double *savr, double *delta, double *r, double *yic, double *ydotic,
                  double *pwk, double *rwm, int *iwm, double cj, double tscale, double sqrtn,
                  double rsqrtn, double epslin, double epscon, double ratemax, int maxit,
                  double stptol, int icnflg, int *icnstr, int *iernew)
{
    int *IWM = iwm - 1;
    double *RWM = rwm - 1;
    int ierr = 0, iersl = 0, ires = 0, iret = 0, m;
    double deltanorm, fnorm = 0.0, fnorm0, oldfnorm, rate, rhok = 0.0, rlx;
    bool iserror, ismaxit;

    int lsoff = IWM[LLSOFF];
    int lrwp = IWM[LLOCWP];
    int liwp = IWM[LLCIWP];
    rate = 1.0;
    rlx = 0.4;
    *iernew = 0;

    for (long i = 0; i < neq; ++i) savr[i] = delta[i];

    dfnrmk(neq, y, t, ydot, savr, r, cj, tscale, wt, sqrtn, rsqrtn, res, &ires, psol, 1,
           &ierr, &fnorm, epslin, &RWM[lrwp], &IWM[liwp], pwk, rpar, ipar);
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
        o_dslvk(neq, y, t, ydot, savr, delta, wt, rwm, iwm, res, &ires, psol, &iersl, cj,
                epslin, sqrtn, rsqrtn, &rhok, rpar, ipar);
        if ((ires != 0) || (iersl != 0)) {
            iserror = true;
            break;
        }
        deltanorm = o_ddwnrm(neq, delta, wt);
        if (deltanorm == 0.0) break;

        oldfnorm = fnorm;
        dlinsk(neq, y, t, ydot, savr, cj, tscale, delta, &deltanorm, wt, sqrtn, rsqrtn, lsoff,
               stptol, &iret, res, &ires, psol, rwm, iwm, rhok, &fnorm, icopt, idy,
               &RWM[lrwp], &IWM[liwp], r, epslin, yic, ydotic, pwk, icnflg, icnstr, rlx,
               rpar, ipar);
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
        for (long i = 0; i < neq; ++i) delta[i] = savr[i];
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

/* src/daskr_new.f90:2004-2173 */
static void ddasik(double t, double *y, double *ydot, long neq, int icopt, const int *idy,
                   o_res_t res, o_jack_t jack, o_psol_t psol, double h, double tscale,
                   double *wt, int *jskip, double *rpar, int *ipar, double *savr,
                   double *delta, double *r, double *y0, double *ydot0, double *pwk,
                   double *rwm, int *iwm, double cj, double uround, double epli,
                   double sqrtn, double rsqrtn, double epscon, double ratemax, double stptol,
                   int jflg, int icnflg, int *icnstr, int *iernls)
{
    (void)uround;
    int *IWM = iwm - 1;
    double *RWM = rwm - 1;
    int neq32 = (int)neq;
    int iernew = 0, ierpj = 0, ires, nj;
    int lrwp = IWM[LLOCWP], liwp = IWM[LLCIWP], maxnit = IWM[LMXNIT], maxnj = IWM[LMXNJ];
    double epslin = epli * epscon;
    bool failed = false;
    nj = 0;

    ires = 0;
    res(&t, y, ydot, &cj, delta, &ires, rpar, ipar);
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
                jack(res, &ires, &neq32, &t, y, ydot, wt, delta, r, &h, &cj, &RWM[lrwp],
                     &IWM[liwp], &ierpj, rpar, ipar);
                if ((ires < 0) || (ierpj != 0)) {
                    failed = true;
                    break;
                }
            }
            *jskip = 0;
            dnsik(t, y, ydot, neq, icopt, idy, res, psol, wt, rpar, ipar, savr, delta, r, y0,
                  ydot0, pwk, rwm, iwm, cj, tscale, sqrtn, rsqrtn, epslin, epscon, ratemax,
                  maxnit, stptol, icnflg, icnstr, &iernew);
            if ((iernew == 1) && (nj < maxnj) && (jflg == 1)) {
                for (long i = 0; i < neq; ++i) delta[i] = savr[i];
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

/* src/daskr.f:2911-3077 */
static void ddasic(double x, double *y, double *yprime, long neq, int icopt, const int *id,
                   o_res_t res, o_jack_t jac, o_psol_t psol, double *h, double tscale,
                   double *wt, int nic, int *idid, double *rpar, int *ipar, double *phi,
                   double *savr, double *delta, double *e, double *yic, double *ypic,
                   double *pwk, double *wm, int *iwm, double uround, double epli,
                   double sqrtn, double rsqrtn, double epconi, double stptol, int jflg,
                   int icnflg, int *icnstr)
{
    const double rhcut = 0.1, ratemx = 0.8;
    int *IWM = iwm - 1;
    int mxnh = IWM[LMXNH];
    int nh, jskip, iernls = 0;
    double cj;
    *idid = 1;
    nh = 1;
    jskip = 0;
    if (nic == 2) jskip = 1;
    o_dcopy(neq, y, phi);
    o_dcopy(neq, yprime, phi + neq);
    if (icopt == 2)
        cj = 0.0;
    else
        cj = 1.0 / (*h);

    for (;;) {
        ddasik(x, y, yprime, neq, icopt, id, res, jac, psol, *h, tscale, wt, &jskip, rpar,
               ipar, savr, delta, e, yic, ypic, pwk, wm, iwm, cj, uround, epli, sqrtn, rsqrtn,
               epconi, ratemx, stptol, jflg, icnflg, icnstr, &iernls);
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
        o_dcopy(neq, phi, y);
        o_dcopy(neq, phi + neq, yprime);
    }
    *idid = -12;
}

/* src/daskr_new.f90:2881-3034 */
static void dnsk(double t, double *y, double *ydot, long neq, o_res_t res, o_psol_t psol,
                 double *wt, double *rpar, int *ipar, double *savr, double *delta, double *e,
                 double *rwm, int *iwm, double cj, double sqrtn, double rsqrtn, double epslin,
                 double epscon, double *s, double confac, double tolnew, int muldel,
                 int maxit, int *ires, int *iersl, int *iernew)
{
    int *IWM = iwm - 1;
    int m = 0;
    double deltanorm, oldnorm = 0.0, rate, rhok;
    bool converged = false;

    for (long i = 0; i < neq; ++i) e[i] = 0.0;

    for (;;) {
        IWM[LNNI] = IWM[LNNI] + 1;
        if (muldel == 1)
            for (long i = 0; i < neq; ++i) delta[i] = delta[i] * confac;
        for (long i = 0; i < neq; ++i) savr[i] = delta[i];

        o_dslvk(neq, y, t, ydot, savr, delta, wt, rwm, iwm, res, ires, psol, iersl, cj,
                epslin, sqrtn, rsqrtn, &rhok, rpar, ipar);
        if ((*ires != 0) || (*iersl != 0)) break;

        for (long i = 0; i < neq; ++i) y[i] = y[i] - delta[i];
        for (long i = 0; i < neq; ++i) e[i] = e[i] - delta[i];
        for (long i = 0; i < neq; ++i) ydot[i] = ydot[i] - cj * delta[i];

        deltanorm = o_ddwnrm(neq, delta, wt);
        if (m == 0) {
            oldnorm = deltanorm;
            if (deltanorm <= tolnew) {
                converged = true;
                break;
            }
        } else {
            rate = std::pow(deltanorm / oldnorm, 1.0 / m);
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
        res(&t, y, ydot, &cj, delta, ires, rpar, ipar);
        IWM[LNRE] = IWM[LNRE] + 1;
        if (*ires < 0) break;
    }

    if (converged) {
        *iernew = 0;
    } else {
        if ((*ires <= -2) || (*iersl < 0))
            *iernew = -1;
        else
            *iernew = 1;
    }
}

/* src/daskr_new.f90:2660-2879 */
static void dnedk(double t, double *y, double *ydot, long neq, o_res_t res, o_jack_t jack,
                  o_psol_t psol, double h, double *wt, int jstart, int *idid, double *rpar,
                  int *ipar, const double *phi, const double *gama, double *savr,
                  double *delta, double *e, double *rwm, int *iwm, double cj, double *cjold,
                  double cjlast, double *s, double uround, double epli, double sqrtn,
                  double rsqrtn, double epscon, int *jcalc, int jflag, int kp1, int nonneg,
                  int ntype, int *iernls)
{
    (void)uround;
    const int muldel = 0, maxit = 4;
    const double xrate = 0.25;
    int *IWM = iwm - 1;
    double *RWM = rwm - 1;
    int neq32 = (int)neq;
    int iernew = 0, ierpj = 0, iersl = 0, iertyp = 0, ires = 0, liwp, lrwp;
    double deltanorm, epslin, temp1 = 0.0, temp2, tolnew;

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
    lrwp = IWM[LLOCWP];
    liwp = IWM[LLCIWP];

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

    for (long i = 0; i < neq; ++i) y[i] = phi[i];
    for (long i = 0; i < neq; ++i) ydot[i] = 0.0;
    for (int j = 2; j <= kp1; ++j) {
        const double *phij = phi + (long)(j - 1) * neq;
        for (long i = 0; i < neq; ++i) y[i] = y[i] + phij[i];
        for (long i = 0; i < neq; ++i) ydot[i] = ydot[i] + gama[j - 1] * phij[i];
    }

    epslin = epli * epscon;
    tolnew = epslin;

    res(&t, y, ydot, &cj, delta, &ires, rpar, ipar);
    IWM[LNRE] = IWM[LNRE] + 1;
    if (ires < 0) goto L380;

    if (*jcalc == -1) {
        *jcalc = 0;
        jack(res, &ires, &neq32, &t, y, ydot, wt, delta, e, &h, &cj, &RWM[lrwp], &IWM[liwp],
             &ierpj, rpar, ipar);
        IWM[LNJE] = IWM[LNJE] + 1;
        *cjold = cj;
        *s = 1e2;
        if (ires < 0) goto L380;
        if (ierpj != 0) goto L380;
    }

    dnsk(t, y, ydot, neq, res, psol, wt, rpar, ipar, savr, delta, e, rwm, iwm, cj, sqrtn,
         rsqrtn, epslin, epscon, s, temp1, tolnew, muldel, maxit, &ires, &iersl, &iernew);

    if (iernew > 0 && *jcalc != 0) {
        *jcalc = -1;
        goto L300;
    }
    if (iernew != 0) goto L380;

    if (nonneg == 0) goto L390;
    for (long i = 0; i < neq; ++i) delta[i] = std::min(y[i], 0.0);
    deltanorm = o_ddwnrm(neq, delta, wt);
    if (deltanorm > epscon) goto L380;
    for (long i = 0; i < neq; ++i) e[i] = e[i] - delta[i];
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

#define PHI(j) (phi + ((long)(j) - 1) * neq)

/* src/daskr_new.f90:149-567 (nls is always dnedk here) */
static void ddstp(double *t, double *y, double *ydot, long neq, o_res_t res, o_jack_t jac,
                  o_psol_t psol, double *h, double *wt, double *vt, int *jstart, int *idid,
                  double *rpar, int *ipar, double *phi, double *savr, double *delta,
                  double *e, double *rwm, int *iwm, double *alpha0_, double *beta0_,
                  double *gama0_, double *psi0_, double *sigma0_, double *cj, double *cjold,
                  double *hold, double *s, double hmin, double uround, double epli,
                  double sqrtn, double rsqrtn, double epscon, int *iphase, int *jcalc,
                  int jflg, int *k, int *kold, int *ns, int nonneg, int ntype)
{
    int *IWM = iwm - 1;
    /* 1-based views of the coefficient arrays */
    double *alpha = alpha0_ - 1, *beta = beta0_ - 1, *gama = gama0_ - 1, *psi = psi0_ - 1,
           *sigma = sigma0_ - 1;
    int iernls = 0, kdiff, km1, knew = 0, kp1, kp2, ncf, nef, nsp1;
    double alpha0, alphas, cjlast, ck, enorm, erk, erkm1 = 0.0, erkm2, erkp1, err, est = 0.0,
           hnew, r = 0.0, temp1, temp2, terk, terkm1 = 0.0, terkm2, terkp1, told;

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

        ck = std::fabs(alpha[kp1] + alphas - alpha0);
        ck = std::max(ck, alpha[kp1]);

        if (kp1 >= nsp1)
            for (int j = nsp1; j <= kp1; ++j)
                for (long i = 0; i < neq; ++i) PHI(j)[i] = beta[j] * PHI(j)[i];

        *t = *t + *h;
        *idid = 1;

        dnedk(*t, y, ydot, neq, res, jac, psol, *h, wt, *jstart, idid, rpar, ipar, phi, gama0_,
              savr, delta, e, rwm, iwm, *cj, cjold, cjlast, s, uround, epli, sqrtn, rsqrtn,
              epscon, jcalc, jflg, kp1, nonneg, ntype, &iernls);

        if (iernls != 0) goto L600;

        /* BLOCK 4: error estimates at orders k, k-1, k-2 */
        enorm = o_ddwnrm(neq, e, vt);
        erk = sigma[*k + 1] * enorm;
        terk = (*k + 1) * erk;
        est = erk;
        knew = *k;

        if (*k == 1) goto L430;
        for (long i = 0; i < neq; ++i) delta[i] = PHI(kp1)[i] + e[i];
        erkm1 = sigma[*k] * o_ddwnrm(neq, delta, vt);
        terkm1 = (*k) * erkm1;
        if (*k > 2) goto L410;
        if (terkm1 <= terk / 2) goto L420;
        goto L430;

    L410:
        for (long i = 0; i < neq; ++i) delta[i] = delta[i] + PHI(*k)[i];
        erkm2 = sigma[*k - 1] * o_ddwnrm(neq, delta, vt);
        terkm2 = (*k - 1) * erkm2;
        if (std::max(terkm1, terkm2) > terk) goto L430;

    L420:
        knew = *k - 1;
        est = erkm1;

    L430:
        err = ck * enorm;
        if (err > 1.0) goto L600;

        /* BLOCK 5: successful step */
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

        for (long i = 0; i < neq; ++i) delta[i] = e[i] - PHI(kp2)[i];
        erkp1 = o_ddwnrm(neq, delta, vt) / (*k + 2);
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
        r = std::pow(2 * est + 1e-4, -1 / temp2);
        if (r < 2.0) goto L555;
        hnew = 2 * (*h);
        goto L560;

    L555:
        if (r > 1.0) goto L560;
        r = std::max(0.5, std::min(0.9, r));
        hnew = (*h) * r;

    L560:
        *h = hnew;

    L575:
        if (*kold == IWM[LMXORD]) goto L585;
        for (long i = 0; i < neq; ++i) PHI(kp2)[i] = e[i];

    L585:
        for (long i = 0; i < neq; ++i) PHI(kp1)[i] = PHI(kp1)[i] + e[i];
        for (int j1 = 2; j1 <= kp1; ++j1) {
            int j = kp1 - j1 + 1;
            for (long i = 0; i < neq; ++i) PHI(j)[i] = PHI(j)[i] + PHI(j + 1)[i];
        }
        *jstart = 1;
        return;

        /* BLOCK 6: unsuccessful step */
    L600:
        *iphase = 1;
        *t = told;
        if (kp1 >= nsp1)
            for (int j = nsp1; j <= kp1; ++j) {
                temp1 = 1 / beta[j];
                for (long i = 0; i < neq; ++i) PHI(j)[i] = temp1 * PHI(j)[i];
            }

        for (int i = 2; i <= kp1; ++i) psi[i - 1] = psi[i] - *h;

        if (iernls == 0) goto L660;
        IWM[LNCFN] = IWM[LNCFN] + 1;

        if (iernls < 0) goto L675;
        ncf = ncf + 1;
        r = 0.25;
        *h = (*h) * r;
        if (ncf < 10 && std::fabs(*h) >= hmin) goto L690;
        if (*idid == 1) *idid = -7;
        if (nef >= 3) *idid = -9;
        goto L675;

    L660:
        nef = nef + 1;
        IWM[LETF] = IWM[LETF] + 1;
        if (nef > 1) goto L665;
        *k = knew;
        temp2 = (double)(*k + 1);
        r = 0.9 * std::pow(2 * est + 1e-4, -1 / temp2);
        r = std::max(0.25, std::min(0.9, r));
        *h = (*h) * r;
        if (std::fabs(*h) >= hmin) goto L690;
        *idid = -6;
        goto L675;

    L665:
        if (nef > 2) goto L670;
        *k = knew;
        r = 0.25;
        *h = r * (*h);
        if (std::fabs(*h) >= hmin) goto L690;
        *idid = -6;
        goto L675;

    L670:
        *k = 1;
        r = 0.25;
        *h = r * (*h);
        if (std::fabs(*h) >= hmin) goto L690;
        *idid = -6;
        goto L675;

    L675:
        o_ddatrp(*t, *t, y, ydot, neq, *k, phi, psi0_);
        *jstart = 1;
        if (*idid >= 0) *idid = -7;
        return;

    L690:
        if (*kold == 0) {
            psi[1] = *h;
            for (long i = 0; i < neq; ++i) PHI(2)[i] = r * PHI(2)[i];
        }
    }
}
#undef PHI

/* callback contract of the root functions, src/daskr.f:908-934 (RT) */
typedef void (*o_rt_t)(const int *neq, const double *t, const double *y, const double *yp, const int *nrt,
                       double *rval, double *rpar, int *ipar);

static inline double sign1(double x) { return std::signbit(x) ? -1.0 : 1.0; } /* SIGN(1.0D0,x) */

/* DROOTS, src/daskr.f:2676-2910.  SAVEd: alpha, x2, imax, last. */
static void droots(int nrt, double hmin, int *jflag, double *x0, double *x1, double *r0, double *r1,
                   double *rx, double *x, int *jroot)
{
    static double alpha, x2;
    static int imax, last;
    int imxold, nxlast = 0;
    double t2, tmax, fracint, fracsub;
    bool zroot, sgnchg, xroot = false;

    if (*jflag == 1) goto L200;
    /* first call: look for sign changes in (x0, x1) */
    imax = 0;
    tmax = 0.0;
    zroot = false;
    for (int i = 0; i < nrt; ++i) {
        if (!(std::fabs(r1[i]) > 0.0)) { zroot = true; continue; }
        if (sign1(r0[i]) == sign1(r1[i])) continue;
        t2 = std::fabs(r1[i] / (r1[i] - r0[i]));
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
    if (nxlast != last)
        alpha = 1.0;
    else if (last != 0)
        alpha = 0.5 * alpha;
    else
        alpha = 2.0 * alpha;
    x2 = *x1 - (*x1 - *x0) * r1[imax - 1] / (r1[imax - 1] - alpha * r0[imax - 1]);
    if (std::fabs(x2 - *x0) < 0.5 * hmin) {
        fracint = std::fabs(*x1 - *x0) / hmin;
        fracsub = (fracint > 5.0) ? 0.1 : 0.5 / fracint;
        x2 = *x0 + fracsub * (*x1 - *x0);
    }
    if (std::fabs(*x1 - x2) < 0.5 * hmin) {
        fracint = std::fabs(*x1 - *x0) / hmin;
        fracsub = (fracint > 5.0) ? 0.1 : 0.5 / fracint;
        x2 = *x1 - fracsub * (*x1 - *x0);
    }
    *jflag = 1;
    *x = x2;
    return;

L200:
    /* re-entry with rx = r(x2): narrow the bracket */
    imxold = imax;
    imax = 0;
    tmax = 0.0;
    zroot = false;
    for (int i = 0; i < nrt; ++i) {
        if (!(std::fabs(rx[i]) > 0.0)) { zroot = true; continue; }
        if (sign1(r0[i]) == sign1(rx[i])) continue;
        t2 = std::fabs(rx[i] / (rx[i] - r0[i]));
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
    if (std::fabs(*x1 - *x0) <= hmin) xroot = true;
    goto L150;

L300:
    *jflag = 2;
    *x = *x1;
    for (int i = 0; i < nrt; ++i) rx[i] = r1[i];
    for (int i = 0; i < nrt; ++i) {
        jroot[i] = 0;
        if (std::fabs(r1[i]) == 0.0) {
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
            if (std::fabs(r1[i]) == 0.0) jroot[i] = (int)(-sign1(r0[i]));
        }
        *jflag = 3;
        return;
    }
    for (int i = 0; i < nrt; ++i) rx[i] = r1[i];
    *x = *x1;
    *jflag = 4;
}

/* DRCHEK, src/daskr.f:2494-2675 */
static void drchek(int job, o_rt_t rt, int nrt, long neq, double tn, double tout, double *y, double *yp,
                   double *phi, double *psi, int kold, double *r0, double *r1, double *rx, int *jroot,
                   int *irt, double uround, int info3, double *rwork, int *iwork, double *rpar, int *ipar)
{
    (void)info3;
    double *RW = rwork - 1;
    int *IW = iwork - 1;
    const int neq32 = (int)neq;
    const double h = psi[0];
    double t1, temp1, temp2, x = 0.0;
    bool zroot;
    int jflag;
    *irt = 0;
    for (int i = 0; i < nrt; ++i) jroot[i] = 0;
    const double hminr = (std::fabs(tn) + std::fabs(h)) * uround * 100.0;

    if (job == 1) {
        /* evaluate R at the initial T (= RWORK(LT0)); check for zero values */
        o_ddatrp(tn, RW[LT0], y, yp, neq, kold, phi, psi);
        rt(&neq32, &RW[LT0], y, yp, &nrt, r0, rpar, ipar);
        IW[LNRTE] = 1;
        zroot = false;
        for (int i = 0; i < nrt; ++i)
            if (std::fabs(r0[i]) == 0.0) zroot = true;
        if (!zroot) return;
        /* R has a zero at T: look at R at T + (small increment) */
        temp2 = std::max(hminr / std::fabs(h), 0.1);
        temp1 = temp2 * h;
        RW[LT0] = RW[LT0] + temp1;
        for (long i = 0; i < neq; ++i) y[i] = y[i] + temp2 * phi[neq + i];
        rt(&neq32, &RW[LT0], y, yp, &nrt, r0, rpar, ipar);
        IW[LNRTE] = IW[LNRTE] + 1;
        zroot = false;
        for (int i = 0; i < nrt; ++i)
            if (std::fabs(r0[i]) == 0.0) zroot = true;
        if (zroot) *irt = -1;
        return;
    }
    if (job == 2) {
        if (IW[LIRFND] != 0) {
            /* a root was found on the previous step: evaluate R0 = R(T0) */
            o_ddatrp(tn, RW[LT0], y, yp, neq, kold, phi, psi);
            rt(&neq32, &RW[LT0], y, yp, &nrt, r0, rpar, ipar);
            IW[LNRTE] = IW[LNRTE] + 1;
            zroot = false;
            for (int i = 0; i < nrt; ++i)
                if (std::fabs(r0[i]) == 0.0) {
                    zroot = true;
                    jroot[i] = 1;
                }
            if (zroot) {
                /* R has a zero at T0: look at R at T0 + (small increment) */
                temp1 = (h >= 0.0) ? std::fabs(hminr) : -std::fabs(hminr);
                RW[LT0] = RW[LT0] + temp1;
                if ((RW[LT0] - tn) * h < 0.0) {
                    o_ddatrp(tn, RW[LT0], y, yp, neq, kold, phi, psi);
                } else {
                    temp2 = temp1 / h;
                    for (long i = 0; i < neq; ++i) y[i] = y[i] + temp2 * phi[neq + i];
                }
                rt(&neq32, &RW[LT0], y, yp, &nrt, r0, rpar, ipar);
                IW[LNRTE] = IW[LNRTE] + 1;
                for (int i = 0; i < nrt; ++i) {
                    if (std::fabs(r0[i]) > 0.0) continue;
                    if (jroot[i] == 1) {   /* R_i has a zero at T0 and also close to T0 */
                        *irt = -2;
                        return;
                    }
                    jroot[i] = (int)(-sign1(r0[i]));
                    *irt = 1;
                }
                if (*irt == 1) return;
            }
        }
        /* R0 has no zero components: proceed to check the relevant interval */
        if (tn == RW[LTLAST]) return;
    }
    /* job == 3 (and the tail of job == 2): set T1 to TN or TOUT, whichever comes first, and get R there */
    if ((tout - tn) * h >= 0.0) {
        t1 = tn;
    } else {
        t1 = tout;
        if ((t1 - RW[LT0]) * h <= 0.0) return;
    }
    o_ddatrp(tn, t1, y, yp, neq, kold, phi, psi);
    rt(&neq32, &t1, y, yp, &nrt, r1, rpar, ipar);
    IW[LNRTE] = IW[LNRTE] + 1;
    jflag = 0;
    for (;;) {
        droots(nrt, hminr, &jflag, &RW[LT0], &t1, r0, r1, rx, &x, jroot);
        if (jflag > 1) break;
        o_ddatrp(tn, x, y, yp, neq, kold, phi, psi);
        rt(&neq32, &x, y, yp, &nrt, rx, rpar, ipar);
        IW[LNRTE] = IW[LNRTE] + 1;
    }
    RW[LT0] = x;
    for (int i = 0; i < nrt; ++i) r0[i] = rx[i];
    if (jflag == 4) return;
    /* found a root: interpolate to X and return */
    o_ddatrp(tn, x, y, yp, neq, kold, phi, psi);
    *irt = 1;
}

extern "C" {

/* SAVEd between calls, src/daskr.f:1429 */
static long s_lid;
static int s_lenid, s_nonneg, s_ncphi;

/* src/daskr.f:1-2493, Krylov option.  lrw/liw are 64-bit here so that CPU
 * baselines at 1024^2 x 20 (LENRW ~ 8e8) stay addressable; every other
 * argument is as in the reference (rt is an o_rt_t passed as void*). */
void o_daskr(o_res_t res, const int *neq_, double *t, double *y, double *yprime,
             const double *tout_, int *info, double *rtol, double *atol,
             int *idid, double *rwork, const long *lrw, int *iwork,
             const long *liw, double *rpar, int *ipar, o_jack_t jac,
             o_psol_t psol, void *rt, const int *nrt_, int *jroot)
{
    o_rt_t rtf = (o_rt_t)rt;
    int irt = 0;
    const long neq = *neq_;
    const int nrt = *nrt_;
    const double tout = *tout_;
    int *IW = iwork - 1;
    double *RW = rwork - 1;
    int *INFO = info - 1;
    int i, mxord, icnflg = 0, index = 1, ier, iret, nwt, nzflg, itemp = 0;
    int nli0, nni0, ncfn0, ncfl0, nwarn, nstd, nnid;
    long lenic, lenpd, lenrw, leniw, lenwp, leniwp, maxl;
    long lsavr, le, lwt, lvt, lphi, lr0, lr1, lrx, lwm, lyic, lypic, lpwk;
    double tn = 0.0, h = 0.0, h0 = 0.0, hmin, hmax, uround, tdist, ypnorm, rh, tstop = 0.0,
           epconi, tscale, r, floatn, rtoli, atoli, tnext;
    bool done;

    if (INFO[1] != 0) goto L100;

    /* ---- initial call: input checks (:1447-1661) ---- */
    for (i = 2; i <= 9; ++i) {
        itemp = i;
        if (INFO[i] != 0 && INFO[i] != 1) goto L701;
    }
    itemp = 10;
    if (INFO[10] < 0 || INFO[10] > 3) goto L701;
    itemp = 11;
    if (INFO[11] < 0 || INFO[11] > 2) goto L701;
    for (i = 12; i <= 17; ++i) {
        itemp = i;
        if (INFO[i] != 0 && INFO[i] != 1) goto L701;
    }
    itemp = 18;
    if (INFO[18] < 0 || INFO[18] > 2) goto L701;
    if (neq <= 0) goto L701;
    if (INFO[12] != 1) goto L701; /* oracle restates the Krylov option only */
    if (nrt < 0) goto L701;

    mxord = 5;
    if (INFO[9] != 0) {
        mxord = IW[LMXORD];
        if (mxord < 1 || mxord > 5) goto L701;
    }
    IW[LMXORD] = mxord;

    icnflg = 0;
    s_nonneg = 0;
    s_lid = LICNS;
    if (INFO[10] == 1) {
        icnflg = 1;
        s_nonneg = 0;
        s_lid = LICNS + neq;
    } else if (INFO[10] == 2) {
        icnflg = 0;
        s_nonneg = 1;
    } else if (INFO[10] == 3) {
        icnflg = 1;
        s_nonneg = 1;
        s_lid = LICNS + neq;
    }

    IW[LMITER] = INFO[12];
    if (INFO[13] == 0) {
        IW[LMAXL] = (int)std::min(5L, neq);
        IW[LKMP] = IW[LMAXL];
        IW[LNRMAX] = 5;
        RW[LEPLI] = 0.05;
    } else {
        if (IW[LMAXL] < 1 || IW[LMAXL] > neq) goto L701;
        if (IW[LKMP] < 1 || IW[LKMP] > IW[LMAXL]) goto L701;
        if (IW[LNRMAX] < 0) goto L701;
        if (RW[LEPLI] <= 0.0 || RW[LEPLI] >= 1.0) goto L701;
    }

    if (INFO[11] != 0) {
        if (INFO[17] == 0) {
            IW[LMXNIT] = 5;
            if (INFO[12] > 0) IW[LMXNIT] = 15;
            IW[LMXNJ] = 6;
            if (INFO[12] > 0) IW[LMXNJ] = 2;
            IW[LMXNH] = 5;
            IW[LLSOFF] = 0;
            RW[LEPIN] = 0.01;
        } else {
            if (IW[LMXNIT] <= 0) goto L701;
            if (IW[LMXNJ] <= 0) goto L701;
            if (IW[LMXNH] <= 0) goto L701;
            if (IW[LLSOFF] < 0 || IW[LLSOFF] > 1) goto L701;
            if (RW[LEPIN] <= 0.0) goto L701;
        }
    }

    /* work-array lengths (:1548-1606) */
    lenic = 0;
    if (INFO[10] == 1 || INFO[10] == 3) lenic = neq;
    s_lenid = 0;
    if (INFO[11] == 1 || INFO[16] == 1) s_lenid = (int)neq;
    s_ncphi = mxord + 1;
    maxl = IW[LMAXL];
    lenwp = IW[LLNWP];
    leniwp = IW[LLNIWP];
    lenpd = (maxl + 3 + std::min(1L, maxl - IW[LKMP])) * neq + (maxl + 3) * maxl + 1 + lenwp;
    lenrw = 60 + 3 * nrt + (mxord + 5) * neq + lenpd;
    leniw = 40 + lenic + s_lenid + leniwp;
    if (INFO[16] != 0) lenrw = lenrw + neq;
    IW[LNIW] = (int)leniw;
    IW[LNRW] = (int)lenrw;
    IW[LNPD] = (int)lenpd;
    IW[LLOCWP] = (int)(lenpd - lenwp + 1);
    if (*lrw < lenrw) goto L701;
    if (*liw < leniw) goto L701;

    if (lenic > 0) {
        for (long ii = 1; ii <= neq; ++ii) {
            int ici = IW[LICNS - 1 + ii];
            if (ici < -2 || ici > 2) goto L701;
        }
        dcnst0(neq, y, &IW[LICNS], &iret);
        if (iret != 0) goto L701;
    }

    index = 1;
    if (s_lenid > 0) {
        index = 0;
        for (long ii = 1; ii <= neq; ++ii) {
            int idi = IW[s_lid - 1 + ii];
            if (idi != 1 && idi != -1) goto L701;
            if (idi == -1) index = 1;
        }
    }

    if (tout == *t) goto L701;
    if (INFO[7] != 0) {
        hmax = RW[LHMAX];
        if (hmax <= 0.0) goto L701;
    }

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
    goto L200;

    /* ---- continuation calls (:1670-1686) ---- */
L100:
    if (INFO[1] == 1) goto L200;
    /* info(1) = -1 left over from an error return, or garbage: fatal */
    goto L701;

    /* ---- all calls (:1696-1736) ---- */
L200:
    IW[LNSTL] = IW[LNST];
    nli0 = IW[LNLI];
    nni0 = IW[LNNI];
    ncfn0 = IW[LNCFN];
    ncfl0 = IW[LNCFL];
    nwarn = 0;
    (void)nli0; (void)nni0; (void)ncfn0; (void)ncfl0; (void)nwarn; (void)nstd; (void)nnid;

    nzflg = 0;
    rtoli = rtol[0];
    atoli = atol[0];
    for (long ii = 1; ii <= neq; ++ii) {
        if (INFO[2] == 1) rtoli = rtol[ii - 1];
        if (INFO[2] == 1) atoli = atol[ii - 1];
        if (rtoli > 0.0 || atoli > 0.0) nzflg = 1;
        if (rtoli < 0.0) goto L701;
        if (atoli < 0.0) goto L701;
        if (INFO[2] == 0) break; /* scalar tolerances: one look is the whole scan */
    }
    if (nzflg == 0) goto L701;

    IW[LLCIWP] = (int)(s_lid + s_lenid);
    lsavr = LDELTA;
    if (INFO[12] != 0) lsavr = LDELTA + neq;
    le = lsavr + neq;
    lwt = le + neq;
    lvt = lwt;
    if (INFO[16] != 0) lvt = lwt + neq;
    lphi = lvt + neq;
    lr0 = lphi + (long)s_ncphi * neq;
    lr1 = lr0 + nrt;
    lrx = lr1 + nrt;
    lwm = lrx + nrt;
    if (INFO[1] == 1) goto L400;

    /* ---- initial call only (:1744-1936) ---- */
    tn = *t;
    *idid = 1;

    o_ddawts(neq, INFO[2], rtol, atol, y, &RW[lwt]);
    o_dinvwt(neq, &RW[lwt], &ier);
    if (ier != 0) goto L701;
    if (INFO[16] != 0)
        for (long ii = 1; ii <= neq; ++ii)
            RW[lvt + ii - 1] = std::max(IW[s_lid + ii - 1], 0) * RW[lwt + ii - 1];

    uround = 2.220446049250313e-16; /* D1MACH(4), src/daux.f:2-11 */
    RW[LROUND] = uround;
    hmin = 4.0 * uround * std::max(std::fabs(*t), std::fabs(tout));

    if (INFO[11] != 0) {
        if (INFO[17] == 0)
            RW[LSTOL] = std::pow(uround, .6667);
        else if (RW[LSTOL] <= 0.0)
            goto L701;
    }

    RW[LEPCON] = 0.33;
    floatn = (double)neq;
    RW[LSQRN] = std::sqrt(floatn);
    RW[LRSQRN] = 1.0 / RW[LSQRN];

    tdist = std::fabs(tout - *t);
    if (tdist < hmin) goto L701;

    if (INFO[8] != 0) {
        h0 = RW[LH];
        if ((tout - *t) * h0 < 0.0) goto L701;
        if (h0 == 0.0) goto L701;
    } else {
        h0 = 0.001 * tdist;
        ypnorm = o_ddwnrm(neq, yprime, &RW[lvt]);
        if (ypnorm > 0.5 / h0) h0 = 0.5 / ypnorm;
        h0 = dsign(h0, tout - *t);
    }
    if (INFO[7] != 0) {
        rh = std::fabs(h0) / RW[LHMAX];
        if (rh > 1.0) h0 = h0 / rh;
    }
    if (INFO[4] != 0) {
        tstop = RW[LTSTOP];
        if ((tstop - *t) * h0 < 0.0) goto L701;
        if ((*t + h0 - tstop) * h0 > 0.0) h0 = tstop - *t;
        if ((tstop - tout) * h0 < 0.0) goto L701;
    }

    if (INFO[11] != 0) {
        /* initial-condition calculation (:1827-1872) */
        nwt = 1;
        epconi = RW[LEPIN] * RW[LEPCON];
        tscale = 0.0;
        if (index == 0) tscale = tdist;
        for (;;) {
            lyic = lwm;
            lypic = lyic + neq;
            lpwk = lypic + neq;
            ddasic(tn, y, yprime, neq, INFO[11], &IW[s_lid], res, jac, psol, &h0, tscale,
                   &RW[lwt], nwt, idid, rpar, ipar, &RW[lphi], &RW[lsavr], &RW[LDELTA],
                   &RW[le], &RW[lyic], &RW[lypic], &RW[lpwk], &RW[lwm], &IW[LIWM], RW[LROUND],
                   RW[LEPLI], RW[LSQRN], RW[LRSQRN], epconi, RW[LSTOL], INFO[15], icnflg,
                   &IW[LICNS]);
            if (*idid < 0) goto L600;
            if (nwt == 2) break;
            nwt = 2;
            o_ddawts(neq, INFO[2], rtol, atol, y, &RW[lwt]);
            o_dinvwt(neq, &RW[lwt], &ier);
            if (ier != 0) goto L701;
        }
        if (INFO[14] == 1) {
            *idid = 4;
            h = h0;
            if (INFO[11] == 1) RW[LHOLD] = h0;
            goto L590;
        }
        o_ddawts(neq, INFO[2], rtol, atol, y, &RW[lwt]);
        o_dinvwt(neq, &RW[lwt], &ier);
        if (ier != 0) goto L701;
        if (INFO[16] != 0)
            for (long ii = 1; ii <= neq; ++ii)
                RW[lvt + ii - 1] = std::max(IW[s_lid + ii - 1], 0) * RW[lwt + ii - 1];

        if (INFO[8] != 0) {
            h0 = RW[LH];
        } else {
            h0 = 0.001 * tdist;
            ypnorm = o_ddwnrm(neq, yprime, &RW[lvt]);
            if (ypnorm > 0.5 / h0) h0 = 0.5 / ypnorm;
            h0 = dsign(h0, tout - *t);
        }
        if (INFO[7] != 0) {
            rh = std::fabs(h0) / RW[LHMAX];
            if (rh > 1.0) h0 = h0 / rh;
        }
        if (INFO[4] != 0) {
            tstop = RW[LTSTOP];
            if ((*t + h0 - tstop) * h0 > 0.0) h0 = tstop - *t;
        }
    }

    /* :1912-1936 */
    h = h0;
    RW[LH] = h;
    for (long ii = 1; ii <= neq; ++ii) {
        RW[lphi + ii - 1] = y[ii - 1];
        RW[lphi + neq + ii - 1] = h * yprime[ii - 1];
    }
    RW[LT0] = *t;
    IW[LIRFND] = 0;
    RW[LPSI] = h;
    RW[LPSI + 1] = 2.0 * h;
    IW[LKOLD] = 1;
    if (nrt != 0) {   /* check for a zero of R near the initial T (:1929-1934) */
        drchek(1, rtf, nrt, neq, *t, tout, y, yprime, &RW[lphi], &RW[LPSI], IW[LKOLD], &RW[lr0], &RW[lr1],
               &RW[lrx], jroot, &irt, RW[LROUND], INFO[3], rwork, iwork, rpar, ipar);
        if (irt < 0) goto L701;
    }
    goto L500;

    /* ---- continuation: stop tests before stepping (:1944-2042) ---- */
L400:
    uround = RW[LROUND];
    done = false;
    tn = RW[LTN];
    h = RW[LH];
    if (nrt != 0) {   /* check for a zero of R near TN (:1949-1963) */
        drchek(2, rtf, nrt, neq, tn, tout, y, yprime, &RW[lphi], &RW[LPSI], IW[LKOLD], &RW[lr0], &RW[lr1],
               &RW[lrx], jroot, &irt, RW[LROUND], INFO[3], rwork, iwork, rpar, ipar);
        if (irt < 0) goto L701;
        if (irt == 1) {
            IW[LIRFND] = 1;
            *idid = 5;
            *t = RW[LT0];
            done = true;
            goto L490;
        }
    }
    if (INFO[7] != 0) {
        rh = std::fabs(h) / RW[LHMAX];
        if (rh > 1.0) h = h / rh;
    }
    if (*t == tout) goto L701;
    if ((*t - tout) * h > 0.0) goto L701;
    if (INFO[4] == 1) goto L430;
    if (INFO[3] == 1) goto L420;
    if ((tn - tout) * h < 0.0) goto L490;
    o_ddatrp(tn, tout, y, yprime, neq, IW[LKOLD], &RW[lphi], &RW[LPSI]);
    *t = tout;
    *idid = 3;
    done = true;
    goto L490;
L420:
    if ((tn - *t) * h <= 0.0) goto L490;
    if ((tn - tout) * h >= 0.0) {
        o_ddatrp(tn, tout, y, yprime, neq, IW[LKOLD], &RW[lphi], &RW[LPSI]);
        *t = tout;
        *idid = 3;
    } else {
        o_ddatrp(tn, tn, y, yprime, neq, IW[LKOLD], &RW[lphi], &RW[LPSI]);
        *t = tn;
        *idid = 1;
    }
    done = true;
    goto L490;
L430:
    tstop = RW[LTSTOP];
    if ((tn - tstop) * h > 0.0) goto L701;
    if ((tstop - tout) * h < 0.0) goto L701;
    if (INFO[3] == 1) goto L440;
    if ((tn - tout) * h < 0.0) goto L450;
    o_ddatrp(tn, tout, y, yprime, neq, IW[LKOLD], &RW[lphi], &RW[LPSI]);
    *t = tout;
    *idid = 3;
    done = true;
    goto L490;
L440:
    if ((tn - *t) * h <= 0.0) goto L450;
    if ((tn - tout) * h >= 0.0) {
        o_ddatrp(tn, tout, y, yprime, neq, IW[LKOLD], &RW[lphi], &RW[LPSI]);
        *t = tout;
        *idid = 3;
    } else {
        o_ddatrp(tn, tn, y, yprime, neq, IW[LKOLD], &RW[lphi], &RW[LPSI]);
        *t = tn;
        *idid = 1;
    }
    done = true;
    goto L490;
L450:
    if (std::fabs(tn - tstop) <= 100.0 * uround * (std::fabs(tn) + std::fabs(h))) {
        o_ddatrp(tn, tstop, y, yprime, neq, IW[LKOLD], &RW[lphi], &RW[LPSI]);
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

    /* ---- step loop (:2053-2168) ---- */
L500:
    if ((IW[LNST] - IW[LNSTL]) >= 500) {
        *idid = -1;
        goto L527;
    }
    /* poor-performance warnings (:2063-2095) only print; omitted */

    o_ddawts(neq, INFO[2], rtol, atol, &RW[lphi], &RW[lwt]);
    o_dinvwt(neq, &RW[lwt], &ier);
    if (ier != 0) {
        *idid = -3;
        goto L527;
    }
    if (INFO[16] != 0)
        for (long ii = 1; ii <= neq; ++ii)
            RW[lvt + ii - 1] = std::max(IW[s_lid + ii - 1], 0) * RW[lwt + ii - 1];

    r = o_ddwnrm(neq, &RW[lphi], &RW[lwt]) * 100.0 * uround;
    if (r > 1.0) {
        if (INFO[2] == 1) {
            for (long ii = 0; ii < neq; ++ii) {
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

    hmin = 4.0 * uround * std::max(std::fabs(tn), std::fabs(tout));
    if (INFO[7] != 0) {
        rh = std::fabs(h) / RW[LHMAX];
        if (rh > 1.0) h = h / rh;
    }

    ddstp(&tn, y, yprime, neq, res, jac, psol, &h, &RW[lwt], &RW[lvt], &INFO[1], idid, rpar,
          ipar, &RW[lphi], &RW[lsavr], &RW[LDELTA], &RW[le], &RW[lwm], &IW[LIWM], &RW[LALPHA],
          &RW[LBETA], &RW[LGAMMA], &RW[LPSI], &RW[LSIGMA], &RW[LCJ], &RW[LCJOLD], &RW[LHOLD],
          &RW[LS], hmin, RW[LROUND], RW[LEPLI], RW[LSQRN], RW[LRSQRN], RW[LEPCON],
          &IW[LPHASE], &IW[LJCALC], INFO[15], &IW[LK], &IW[LKOLD], &IW[LNS], s_nonneg,
          INFO[12]);

L527:
    if (*idid < 0) goto L600;

    /* ---- successful step: root check (:2177-2189), then stop tests (:2191-2245) ---- */
    if (nrt != 0) {
        drchek(3, rtf, nrt, neq, tn, tout, y, yprime, &RW[lphi], &RW[LPSI], IW[LKOLD], &RW[lr0], &RW[lr1],
               &RW[lrx], jroot, &irt, RW[LROUND], INFO[3], rwork, iwork, rpar, ipar);
        if (irt == 1) {
            IW[LIRFND] = 1;
            *idid = 5;
            *t = RW[LT0];
            goto L590;
        }
    }
    if (INFO[4] == 0) {
        if ((tn - tout) * h >= 0.0) {
            o_ddatrp(tn, tout, y, yprime, neq, IW[LKOLD], &RW[lphi], &RW[LPSI]);
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
    if (INFO[3] == 0) {
        if (std::fabs(tn - tstop) <= 100.0 * uround * (std::fabs(tn) + std::fabs(h))) {
            o_ddatrp(tn, tstop, y, yprime, neq, IW[LKOLD], &RW[lphi], &RW[LPSI]);
            *t = tstop;
            *idid = 2;
            goto L590;
        }
        if ((tn - tout) * h >= 0.0) {
            o_ddatrp(tn, tout, y, yprime, neq, IW[LKOLD], &RW[lphi], &RW[LPSI]);
            *t = tout;
            *idid = 3;
            goto L590;
        }
        tnext = tn + h;
        if ((tnext - tstop) * h <= 0.0) goto L500;
        h = tstop - tn;
        goto L500;
    }
    if (std::fabs(tn - tstop) <= 100.0 * uround * (std::fabs(tn) + std::fabs(h))) {
        o_ddatrp(tn, tstop, y, yprime, neq, IW[LKOLD], &RW[lphi], &RW[LPSI]);
        *t = tstop;
        *idid = 2;
        goto L590;
    }
    if ((tn - tout) * h >= 0.0) {
        o_ddatrp(tn, tout, y, yprime, neq, IW[LKOLD], &RW[lphi], &RW[LPSI]);
        *t = tout;
        *idid = 3;
        goto L590;
    }
    *t = tn;
    *idid = 1;

    /* ---- successful returns (:2252-2256) ---- */
L590:
    RW[LTN] = tn;
    RW[LTLAST] = *t;
    RW[LH] = h;
    return;

    /* ---- unsuccessful returns other than illegal input (:2263-2386) ---- */
L600:
    INFO[1] = -1;
    *t = tn;
    RW[LTN] = tn;
    RW[LH] = h;
    return;

    /* ---- illegal input (:2390-2490): idid = -33, info(1) = -1 ---- */
L701:
    (void)itemp;
    INFO[1] = -1;
    *idid = -33;
    return;
}

/* ------------------------------------------------------------------ */
/* whole-run helpers                                                   */
/* ------------------------------------------------------------------ */

/* Optional per-step log for the CPU baseline timing (bench.py): after every BDF step taken in
 * max_steps mode the run helpers append (wall seconds since the log was armed, NLI, NST). */
static double *g_steplog = nullptr;
static long g_steplog_cap = 0, g_steplog_n = 0;
static double g_steplog_t0 = 0.0;
static double wall_now()
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
void o_set_steplog(double *buf, long cap)
{
    g_steplog = buf;
    g_steplog_cap = cap;
    g_steplog_n = 0;
    g_steplog_t0 = wall_now();
}
long o_steplog_count(void) { return g_steplog_n; }
static void steplog_push(const int *iwork)
{
    if (!g_steplog || g_steplog_n >= g_steplog_cap) return;
    g_steplog[3 * g_steplog_n + 0] = wall_now() - g_steplog_t0;
    g_steplog[3 * g_steplog_n + 1] = iwork[LNLI - 1];
    g_steplog[3 * g_steplog_n + 2] = iwork[LNST - 1];
    ++g_steplog_n;
}

/* Drive DASKR over a list of output times the way example_web.f90:884-954
 * does: first call computes consistent initial values (info(11)=1,
 * info(14)=1 -> idid=4), then info(11)=0 and integration restarts from the
 * corrected values.  nrt=0 (rootfinding does not alter the step sequence). */
static int run_web_impl(int nxg, int nyg, int np, int mx, int my, int jpre, double rtol_, double atol_,
                        int nout, const double *touts, double *yout, int *counters,
                        double *rhead, long max_steps, double *tfinal);

int o_run_web(int np, int mx, int my, int jpre, double rtol_, double atol_,
              int nout, const double *touts, double *yout, int *counters,
              double *rhead, long max_steps, double *tfinal)
{
    return run_web_impl(0, 0, np, mx, my, jpre, rtol_, atol_, nout, touts, yout, counters, rhead, max_steps, tfinal);
}

int o_run_web_bg(int nxg, int nyg, int np, int mx, int my, int jpre, double rtol_, double atol_,
                 int nout, const double *touts, double *yout, int *counters,
                 double *rhead, long max_steps, double *tfinal)
{
    return run_web_impl(nxg, nyg, np, mx, my, jpre, rtol_, atol_, nout, touts, yout, counters, rhead, max_steps, tfinal);
}

static int run_web_impl(int nxg, int nyg, int np, int mx, int my, int jpre, double rtol_, double atol_,
                        int nout, const double *touts, double *yout, int *counters,
                        double *rhead, long max_steps, double *tfinal)
{
    o_web_setpar(np, mx, my);
    const int ns = 2 * np;
    const long neq = (long)ns * mx * my;
    int neq32 = (int)neq;
    int info[20];
    std::memset(info, 0, sizeof(info));
    info[10] = 1; /* info(11) */
    info[11] = 1; /* info(12) Krylov */
    info[13] = 1; /* info(14) */
    info[14] = 1; /* info(15) */
    info[15] = 1; /* info(16) */
    const int jbg = (nxg > 0) ? 1 : 0;
    int ipar[2] = {jpre, jbg};
    std::vector<int> iwork(40 + 2 * neq + 16, 0);
    if (jbg == 0) o_setup_rbdpre(mx, my, ns, np, 40, iwork.data());
    else o_setup_rbgpre(mx, my, ns, np, nxg, nyg, 40, iwork.data());
    long lenwp = iwork[26];
    long lrw = 104 + 19 * neq + lenwp + 64;
    long liw = (long)iwork.size();
    std::vector<double> rwork(lrw, 0.0), c(neq), cdot(neq), rpar(neq);
    double rtol = rtol_, atol = atol_;
    double t = 0.0, tout = touts[0];
    int idid = 0, nrt = 0, jroot = 0;

    /* flat predator guess: the example uses 1e5 for np = 1 (example_web.f90:784), i.e. ee * (prey level 10) * np;
       scaled with np so that Newton starts near the non-trivial predator branch for np > 1 as well */
    o_web_cinit(c.data(), cdot.data(), 1e5 * np, rpar.data());

    /* initial-condition call */
    o_daskr(o_web_res, &neq32, &t, c.data(), cdot.data(), &tout, info, &rtol, &atol, &idid,
            rwork.data(), &lrw, iwork.data(), &liw, rpar.data(), ipar, o_web_jac, o_web_psol,
            nullptr, &nrt, &jroot);
    if (idid < 0) {
        *tfinal = t;
        return idid;
    }
    info[10] = 0;
    if (max_steps > 0) info[2] = 1; /* info(3): return after every step */

    long steps = 0;
    for (int iout = 0; iout < nout; ++iout) {
        tout = touts[iout];
        for (;;) {
            o_daskr(o_web_res, &neq32, &t, c.data(), cdot.data(), &tout, info, &rtol, &atol,
                    &idid, rwork.data(), &lrw, iwork.data(), &liw, rpar.data(), ipar,
                    o_web_jac, o_web_psol, nullptr, &nrt, &jroot);
            if (idid < 0) break;
            if (max_steps > 0 && idid == 1) {
                ++steps;
                steplog_push(iwork.data());
                if (steps >= max_steps) break;
                continue;
            }
            break;
        }
        if (yout) std::memcpy(yout + (long)iout * neq, c.data(), sizeof(double) * neq);
        if (counters) std::memcpy(counters + (long)iout * 40, iwork.data(), sizeof(int) * 40);
        if (rhead) std::memcpy(rhead + (long)iout * 60, rwork.data(), sizeof(double) * 60);
        *tfinal = t;
        if (idid < 0) return idid;
        if (max_steps > 0 && steps >= max_steps) return idid;
    }
    return idid;
}

/* Heat problem as example_heat.f90:174-240 but with the rbdpre (ns=1)
 * diagonal preconditioner of BASELINE config 3; nrt=0. */
int o_run_heat(int m, double rtol_, double atol_, int nout, const double *touts,
               double *yout, int *counters, double *rhead, long max_steps,
               double *tfinal)
{
    o_heat_setpar(m);
    const long neq = (long)(m + 2) * (m + 2);
    int neq32 = (int)neq;
    int info[20];
    std::memset(info, 0, sizeof(info));
    info[11] = 1; /* info(12) */
    info[14] = 1; /* info(15) */
    int ipar[4] = {0, 0, neq32, m};
    std::vector<int> iwork(40 + neq + 16, 0);
    o_setup_rbdpre(m + 2, m + 2, 1, 1, 0, iwork.data());
    long lenwp = iwork[26];
    long lrw = 107 + 18 * neq + lenwp + 64;
    long liw = (long)iwork.size();
    std::vector<double> rwork(lrw, 0.0), u(neq), udot(neq), rpar(2);
    rpar[0] = 1.0 / (m + 1);
    rpar[1] = 1.0 / (rpar[0] * rpar[0]);
    double rtol = rtol_, atol = atol_;
    double t = 0.0, tout;
    int idid = 0, nrt = 0, jroot = 0;
    o_heat_uinit(u.data(), udot.data());
    if (max_steps > 0) info[2] = 1;

    long steps = 0;
    for (int iout = 0; iout < nout; ++iout) {
        tout = touts[iout];
        for (;;) {
            o_daskr(o_heat_res, &neq32, &t, u.data(), udot.data(), &tout, info, &rtol, &atol,
                    &idid, rwork.data(), &lrw, iwork.data(), &liw, rpar.data(), ipar,
                    o_heat_jac, o_heat_psol, nullptr, &nrt, &jroot);
            if (idid < 0) break;
            if (max_steps > 0 && idid == 1) {
                ++steps;
                steplog_push(iwork.data());
                if (steps >= max_steps) break;
                continue;
            }
            break;
        }
        if (yout) std::memcpy(yout + (long)iout * neq, u.data(), sizeof(double) * neq);
        if (counters) std::memcpy(counters + (long)iout * 40, iwork.data(), sizeof(int) * 40);
        if (rhead) std::memcpy(rhead + (long)iout * 60, rwork.data(), sizeof(double) * 60);
        *tfinal = t;
        if (idid < 0) return idid;
        if (max_steps > 0 && steps >= max_steps) return idid;
    }
    return idid;
}

} /* extern "C" */
