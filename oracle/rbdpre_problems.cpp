/*
 * oracle/rbdpre_problems.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restatement of the reaction-based block-diagonal preconditioner
 * (src/daskr_rbdpre.f90:95-279) and of the two user problems that drive the
 * Krylov path: the food web (example/example_web.f90:5-60, 102-148, 190-299,
 * 301-391, 393-598) and the heat equation (example/example_heat.f90:19-89).
 * Like the reference, both keep their parameters in file-scope state
 * ("module variables", thread unsafe).
 */
#include "oracle.h"
#include <cmath>
#include <vector>
#include <algorithm>

extern "C" {

/* ------------------------------------------------------------------ */
/* daskr_rbdpre module variables, src/daskr_rbdpre.f90:95-96           */
/* ------------------------------------------------------------------ */
static int mp, mpd, mpsq, meshx, meshy, mxmp;
static double srur;

/* src/daskr_rbdpre.f90:102-156.  lid is the 1-based iwork offset of the ID
 * array (iwork(lid+1 ...)); iwork(27:28) get LENWP / LENIWP.  The reference
 * stores these in default INTEGERs; sizes beyond 2^31-1 are outside the
 * oracle's range. */
void o_setup_rbdpre(int mx, int my, int ns, int nsd, int lid, int *iwork)
{
    int *IW = iwork - 1;
    srur = std::sqrt(2.220446049250313e-16); /* sqrt(epsilon(one)) = 2^-26 */
    mp = ns;
    mpd = nsd;
    mpsq = ns * ns;
    meshx = mx;
    meshy = my;
    mxmp = meshx * mp;

    IW[27] = mpsq * meshx * meshy;
    IW[28] = mp * meshx * meshy;

    if (lid == 0) return;
    long i0 = lid;
    for (int jy = 1; jy <= my; ++jy)
        for (int jx = 1; jx <= mx; ++jx) {
            for (int i = 1; i <= mpd; ++i) IW[i0 + i] = 1;
            for (int i = mpd + 1; i <= mp; ++i) IW[i0 + i] = -1;
            i0 = i0 + mp;
        }
}

/* src/daskr_rbdpre.f90:158-251 */
void o_jac_rbdpre(double t, double *u, const double *r0, o_rblock_t rblock,
                  double *r1, const double *rewt, double cj, double *bd,
                  int *ipbd, int *ierr)
{
    const double dfac = 1e-2;
    long ibd = 0, j0 = 0;
    for (int jy = 1; jy <= meshy; ++jy) {
        for (int jx = 1; jx <= meshx; ++jx) {
            for (int js = 1; js <= mp; ++js) {
                long j = j0 + js; /* 1-based */
                double uj = u[j - 1];
                double del = std::max(srur * std::fabs(uj), dfac / rewt[j - 1]);
                u[j - 1] = u[j - 1] + del;
                double fac = -1.0 / del;
                rblock(&t, &jx, &jy, &u[j0], r1);
                for (int i = 1; i <= mp; ++i)
                    bd[ibd + i - 1] = (r1[i - 1] - r0[j0 + i - 1]) * fac;
                u[j - 1] = uj;
                ibd = ibd + mp;
            }
            j0 = j0 + mp;
        }
    }

    /* add cj*I_d and factor, stopping at the first singular block (:236-249) */
    ibd = 1;
    long iip = 1;
    long nblk = (long)meshx * meshy;
    for (long jb = 1; jb <= nblk; ++jb) {
        long idiag = ibd;
        for (int i = 1; i <= mp; ++i) {
            if (i <= mpd) bd[idiag - 1] = bd[idiag - 1] + cj;
            idiag = idiag + (mp + 1);
        }
        o_dgefa(&bd[ibd - 1], mp, mp, &ipbd[iip - 1], ierr);
        if (*ierr != 0) break;
        ibd = ibd + mpsq;
        iip = iip + mp;
    }
}

/* src/daskr_rbdpre.f90:253-279 */
void o_psol_rbdpre(double *b, const double *bd, const int *ipbd)
{
    long ib = 1, ibd = 1;
    for (int jy = 1; jy <= meshy; ++jy)
        for (int jx = 1; jx <= meshx; ++jx) {
            o_dgesl(&bd[ibd - 1], mp, mp, &ipbd[ib - 1], &b[ib - 1], 0);
            ib = ib + mp;
            ibd = ibd + mpsq;
        }
}

/* ------------------------------------------------------------------ */
/* block-grouped variant: src/daskr_rbgpre.f90                          */
/* ------------------------------------------------------------------ */
static int g_ngx, g_ngy, g_ngrp;
static std::vector<int> g_jgx, g_jgy, g_jigx, g_jigy, g_jxr, g_jyr; /* 1-based, element 0 unused */

/* src/daskr_rbgpre.f90:160-193 (set_grouping) */
static void set_grouping(int m, int ng, std::vector<int> &jg, std::vector<int> &jig, std::vector<int> &jr)
{
    int mper = m / ng;
    for (int ig = 1; ig <= ng; ++ig) jg[ig] = 1 + (ig - 1) * mper;
    jg[ng + 1] = m + 1;
    int ngm1 = ng - 1;
    int len1 = ngm1 * mper;
    for (int j = 1; j <= len1; ++j) jig[j] = 1 + (j - 1) / mper;
    len1 = len1 + 1;
    for (int j = len1; j <= m; ++j) jig[j] = ng;
    for (int ig = 1; ig <= ngm1; ++ig) jr[ig] = (int)(0.5 + ((double)ig - 0.5) * mper);
    jr[ng] = (int)(0.5 * (1 + ngm1 * mper + m));
}

/* src/daskr_rbgpre.f90:107-158 */
void o_setup_rbgpre(int mx, int my, int ns, int nsd, int nxg, int nyg, int lid, int *iwork)
{
    int *IW = iwork - 1;
    srur = std::sqrt(2.220446049250313e-16);
    mp = ns;
    mpd = nsd;
    mpsq = ns * ns;
    meshx = mx;
    meshy = my;
    mxmp = meshx * mp;
    g_ngx = nxg;
    g_ngy = nyg;
    g_ngrp = nxg * nyg;
    int maxm = std::max(meshx, meshy);
    g_jgx.assign(maxm + 2, 0); g_jgy.assign(maxm + 2, 0);
    g_jigx.assign(maxm + 1, 0); g_jigy.assign(maxm + 1, 0);
    g_jxr.assign(maxm + 1, 0); g_jyr.assign(maxm + 1, 0);
    set_grouping(meshx, g_ngx, g_jgx, g_jigx, g_jxr);
    set_grouping(meshy, g_ngy, g_jgy, g_jigy, g_jyr);
    IW[27] = mpsq * g_ngrp;
    IW[28] = mp * g_ngrp;
    if (lid == 0) return;
    long i0 = lid;
    for (int jy = 1; jy <= my; ++jy)
        for (int jx = 1; jx <= mx; ++jx) {
            for (int i = 1; i <= mpd; ++i) IW[i0 + i] = 1;
            for (int i = mpd + 1; i <= mp; ++i) IW[i0 + i] = -1;
            i0 = i0 + mp;
        }
}

/* src/daskr_rbgpre.f90:195-276: one block per group, evaluated at the group's representative point */
void o_jac_rbgpre(double t, double *u, const double *r0, o_rblock_t rblock, double *r1,
                  const double *rewt, double cj, double *bd, int *ipbd, int *ierr)
{
    const double dfac = 1e-2;
    long ibd = 0;
    for (int igy = 1; igy <= g_ngy; ++igy) {
        int jy = g_jyr[igy];
        long j00 = (long)(jy - 1) * mxmp;
        for (int igx = 1; igx <= g_ngx; ++igx) {
            int jx = g_jxr[igx];
            long j0 = j00 + (long)(jx - 1) * mp;
            for (int js = 1; js <= mp; ++js) {
                long j = j0 + js;
                double uj = u[j - 1];
                double del = std::max(srur * std::fabs(uj), dfac / rewt[j - 1]);
                u[j - 1] = u[j - 1] + del;
                double fac = -1.0 / del;
                rblock(&t, &jx, &jy, &u[j0], r1);
                for (int i = 1; i <= mp; ++i) bd[ibd + i - 1] = (r1[i - 1] - r0[j0 + i - 1]) * fac;
                u[j - 1] = uj;
                ibd = ibd + mp;
            }
        }
    }
    ibd = 1;
    long iip = 1;
    for (int ig = 1; ig <= g_ngrp; ++ig) {
        long idiag = ibd;
        for (int i = 1; i <= mp; ++i) {
            if (i <= mpd) bd[idiag - 1] = bd[idiag - 1] + cj;
            idiag = idiag + (mp + 1);
        }
        o_dgefa(&bd[ibd - 1], mp, mp, &ipbd[iip - 1], ierr);
        if (*ierr != 0) break;
        ibd = ibd + mpsq;
        iip = iip + mp;
    }
}

/* src/daskr_rbgpre.f90:278-304: every mesh point is solved with its group's factors */
void o_psol_rbgpre(double *b, const double *bd, const int *ipbd)
{
    long ib = 1;
    for (int jy = 1; jy <= meshy; ++jy) {
        int igy = g_jigy[jy];
        int ig0 = (igy - 1) * g_ngx;
        for (int jx = 1; jx <= meshx; ++jx) {
            int igx = g_jigx[jx];
            int igm1 = igx - 1 + ig0;
            long ibd = 1 + (long)igm1 * mpsq;
            long iip = 1 + (long)igm1 * mp;
            o_dgesl(&bd[ibd - 1], mp, mp, &ipbd[iip - 1], &b[ib - 1], 0);
            ib = ib + mp;
        }
    }
}

/* ------------------------------------------------------------------ */
/* food web: module web_par, example/example_web.f90:5-60              */
/* ------------------------------------------------------------------ */
static int w_np, w_ns, w_mx, w_my, w_mxns;
static double w_aa, w_ee, w_gg, w_bb, w_dprey, w_dpred, w_alpha, w_beta, w_ax, w_ay, w_dx, w_dy;
static std::vector<double> w_acoef, w_bcoef, w_cox, w_coy, w_diff;
static double w_pi;
static int w_jyoff = 0; /* o_web_window: global row of window row 1 minus 1 (0 for the whole mesh) */
#define ACOEF(i, j) w_acoef[((j) - 1) * w_ns + ((i) - 1)]

/* example/example_web.f90:19-58, with np / mx / my as run-time values
 * (the reference fixes them as parameters at :11) */
void o_web_setpar(int np, int mx, int my)
{
    w_np = np;
    w_ns = 2 * np;
    w_mx = mx;
    w_my = my;
    w_jyoff = 0;
    w_mxns = mx * w_ns;
    w_pi = 4 * std::atan(1.0);
    w_acoef.assign((size_t)w_ns * w_ns, 0.0);
    w_bcoef.assign(w_ns, 0.0);
    w_cox.assign(w_ns, 0.0);
    w_coy.assign(w_ns, 0.0);
    w_diff.assign(w_ns, 0.0);

    w_ax = 1.0;
    w_ay = 1.0;
    w_aa = 1.0;
    w_ee = 1e4;
    w_gg = 0.5e-6;
    w_bb = 1.0;
    w_dprey = 1.0;
    w_dpred = 0.05;
    w_alpha = 5e1;
    w_beta = 3e2;

    w_dx = w_ax / (mx - 1);
    w_dy = w_ay / (my - 1);

    for (int j = 1; j <= np; ++j) {
        for (int i = 1; i <= np; ++i) {
            ACOEF(np + i, j) = w_ee;
            ACOEF(i, np + j) = -w_gg;
        }
        ACOEF(j, j) = -w_aa;
        ACOEF(np + j, np + j) = -w_aa;
        w_bcoef[j - 1] = w_bb;
        w_bcoef[np + j - 1] = -w_bb;
        w_diff[j - 1] = w_dprey;
        w_diff[np + j - 1] = w_dpred;
    }
    for (int i = 1; i <= w_ns; ++i) {
        w_cox[i - 1] = w_diff[i - 1] / (w_dx * w_dx);
        w_coy[i - 1] = w_diff[i - 1] / (w_dy * w_dy);
    }
}

/* Test sampling aid (not in the reference): after o_web_setpar(np, mx, my_global), restrict the row loops of
 * f / res / cinit to a window of `my_window` rows whose first row is global row jy_off + 1.  Mesh spacing and
 * coefficients stay those of the whole mesh; the window's first and last rows see mirrored neighbours (wrong unless
 * they are the global boundary), so callers compare interior window rows only.  Lets the full-size GPU kernels be
 * checked against the restated arithmetic on a few rows of a 4096 x 4096 mesh. */
void o_web_window(int my_window, int jy_off)
{
    w_my = my_window;
    w_jyoff = jy_off;
}

/* example/example_web.f90:271-299 */
void o_web_rates(const double *t, const int *jx, const int *jy, const double *c, double *rxy)
{
    (void)t;
    const int ns = w_ns;
    double y = (*jy - 1 + w_jyoff) * w_dy;
    double x = (*jx - 1) * w_dx;
    double fac = 1.0 + w_alpha * x * y + w_beta * std::sin(4 * w_pi * x) * std::sin(4 * w_pi * y);

    for (int k = 0; k < ns; ++k) rxy[k] = 0.0;
    for (int i = 1; i <= ns; ++i)
        for (int k = 1; k <= ns; ++k) rxy[k - 1] = rxy[k - 1] + c[i - 1] * ACOEF(k, i);

    for (int i = 0; i < ns; ++i) rxy[i] = c[i] * (w_bcoef[i] * fac + rxy[i]);
}

/* example/example_web.f90:226-269 (c, cdot, rpar 0-based here; ic is 1-based) */
void o_web_f(double t, const double *c, double *cdot, double *rpar)
{
    const int ns = w_ns, mx = w_mx, my = w_my, mxns = w_mxns;
    const double *C = c - 1;
    double *CD = cdot - 1, *RP = rpar - 1;
    for (int jy = 1; jy <= my; ++jy) {
        long iyoff = (long)mxns * (jy - 1);
        long idyu = mxns;
        if (jy == my) idyu = -mxns;
        long idyl = mxns;
        if (jy == 1) idyl = -mxns;
        for (int jx = 1; jx <= mx; ++jx) {
            long ic = iyoff + (long)ns * (jx - 1) + 1;
            o_web_rates(&t, &jx, &jy, &C[ic], &RP[ic]);
            long idxu = ns;
            if (jx == mx) idxu = -ns;
            long idxl = ns;
            if (jx == 1) idxl = -ns;
            for (int i = 1; i <= ns; ++i) {
                long ici = ic + i - 1;
                double dcyli = C[ici] - C[ici - idyl];
                double dcyui = C[ici + idyu] - C[ici];
                double dcxli = C[ici] - C[ici - idxl];
                double dcxui = C[ici + idxu] - C[ici];
                CD[ici] = w_coy[i - 1] * (dcyui - dcyli) + w_cox[i - 1] * (dcxui - dcxli) + RP[ici];
            }
        }
    }
}

/* example/example_web.f90:190-224 */
void o_web_res(const double *t, const double *c, const double *cdot, const double *cj,
               double *delta, int *ires, double *rpar, int *ipar)
{
    (void)cj;
    (void)ires;
    (void)ipar;
    const int ns = w_ns, np = w_np, mx = w_mx, my = w_my, mxns = w_mxns;
    o_web_f(*t, c, delta, rpar);
    for (int jy = 1; jy <= my; ++jy) {
        long iyoff = (long)mxns * (jy - 1);
        for (int jx = 1; jx <= mx; ++jx) {
            long ic0 = iyoff + (long)ns * (jx - 1);
            for (int i = 1; i <= ns; ++i) {
                long ici = ic0 + i - 1; /* 0-based */
                if (i > np)
                    delta[ici] = -delta[ici];
                else
                    delta[ici] = cdot[ici] - delta[ici];
            }
        }
    }
}

/* example/example_web.f90:102-148 */
void o_web_cinit(double *c, double *cdot, double pred_ic, double *rpar)
{
    const int ns = w_ns, np = w_np, mx = w_mx, my = w_my, mxns = w_mxns;
    int npp1 = np + 1;
    for (int jy = 1; jy <= my; ++jy) {
        double y = (jy - 1) * w_dy;
        double argy = 16 * y * y * (w_ay - y) * (w_ay - y);
        long iyoff = (long)mxns * (jy - 1);
        for (int jx = 1; jx <= mx; ++jx) {
            double x = (jx - 1) * w_dx;
            double argx = 16 * x * x * (w_ax - x) * (w_ax - x);
            long ioff = iyoff + (long)ns * (jx - 1);
            for (int i = 1; i <= np; ++i) c[ioff + i - 1] = 10.0 + i * argx * argy;
            for (int i = npp1; i <= ns; ++i) c[ioff + i - 1] = pred_ic;
        }
    }
    o_web_f(0.0, c, cdot, rpar);
    for (int jy = 1; jy <= my; ++jy) {
        long iyoff = (long)mxns * (jy - 1);
        for (int jx = 1; jx <= mx; ++jx) {
            long ioff = iyoff + (long)ns * (jx - 1);
            for (int i = npp1; i <= ns; ++i) cdot[ioff + i - 1] = 0.0;
        }
    }
}

/* example/example_web.f90:301-344; jbg=ipar(2) must be 0 (no grouping) */
void o_web_jac(o_res_t res, int *ires, const int *neq, const double *t, double *c,
               double *cdot, const double *rewt, double *savr, double *wk,
               const double *h, const double *cj, double *rwp, int *iwp,
               int *ierr, double *rpar, int *ipar)
{
    (void)res; (void)ires; (void)neq; (void)cdot; (void)savr; (void)h;
    int jbg = ipar[1]; /* example_web.f90:337-342 */
    if (jbg == 0) o_jac_rbdpre(*t, c, rpar, o_web_rates, wk, rewt, *cj, rwp, iwp, ierr);
    else o_jac_rbgpre(*t, c, rpar, o_web_rates, wk, rewt, *cj, rwp, iwp, ierr);
}

/* example/example_web.f90:393-575 -- itmax=5 lexicographic Gauss-Seidel sweeps
 * on A_S = I - hl0*J_d.  x and z are 1-based below. */
void o_web_gauss_seidel(long n, double hl0, double *z0, double *x0)
{
    const int itmax = 5;
    const int ns = w_ns, mx = w_mx, my = w_my;
    const long mxns = w_mxns;
    double *z = z0 - 1, *x = x0 - 1;
    std::vector<double> beta1(ns + 1), gamma1(ns + 1), beta2(ns + 1), gamma2(ns + 1), dinv(ns + 1);

    for (int i = 1; i <= ns; ++i) {
        double elamda = 1.0 / (1.0 + 2 * hl0 * (w_cox[i - 1] + w_coy[i - 1]));
        beta1[i] = hl0 * w_cox[i - 1] * elamda;
        beta2[i] = 2 * beta1[i];
        gamma1[i] = hl0 * w_coy[i - 1] * elamda;
        gamma2[i] = 2 * gamma1[i];
        dinv[i] = elamda;
    }
    for (long i = 1; i <= n; ++i) x[i] = 0.0;

    for (int jy = 1; jy <= my; ++jy) {
        long iyoff = mxns * (jy - 1);
        for (int jx = 1; jx <= mx; ++jx) {
            long ic = iyoff + (long)ns * (jx - 1);
            for (int i = 1; i <= ns; ++i) {
                long ici = ic + i;
                x[ici] = dinv[i] * z[ici];
                z[ici] = 0.0;
            }
        }
    }

    for (int iter = 1; iter <= itmax; ++iter) {
        /* (D-inverse)*U*X */
        if (iter > 1) {
            for (int jy = 1; jy <= my; ++jy) {
                long iyoff = mxns * (jy - 1);
                const std::vector<double> &gam = (jy == 1) ? gamma2 : gamma1;
                bool lasty = (jy == my);
                for (int jx = 1; jx <= mx; ++jx) {
                    long ic = iyoff + (long)ns * (jx - 1);
                    for (int i = 1; i <= ns; ++i) {
                        long ici = ic + i;
                        if (jx == 1) {
                            if (!lasty)
                                x[ici] = beta2[i] * x[ici + ns] + gam[i] * x[ici + mxns];
                            else
                                x[ici] = beta2[i] * x[ici + ns];
                        } else if (jx < mx) {
                            if (!lasty)
                                x[ici] = beta1[i] * x[ici + ns] + gam[i] * x[ici + mxns];
                            else
                                x[ici] = beta1[i] * x[ici + ns];
                        } else {
                            if (!lasty)
                                x[ici] = gam[i] * x[ici + mxns];
                            else
                                x[ici] = 0.0;
                        }
                    }
                }
            }
        }
        /* [(I - (D-inverse)*L)]-inverse * X */
        for (int jy = 1; jy <= my; ++jy) {
            long iyoff = mxns * (jy - 1);
            const std::vector<double> &gam = (jy == my) ? gamma2 : gamma1;
            for (int jx = 1; jx <= mx; ++jx) {
                if (jy == 1 && jx == 1) continue;
                long ic = iyoff + (long)ns * (jx - 1);
                const std::vector<double> &bet = (jx == mx) ? beta2 : beta1;
                for (int i = 1; i <= ns; ++i) {
                    long ici = ic + i;
                    if (jy == 1)
                        x[ici] = x[ici] + bet[i] * x[ici - ns];
                    else if (jx == 1)
                        x[ici] = x[ici] + gam[i] * x[ici - mxns];
                    else
                        x[ici] = (x[ici] + bet[i] * x[ici - ns]) + gam[i] * x[ici - mxns];
                }
            }
        }
        for (long i = 1; i <= n; ++i) z[i] = z[i] + x[i];
    }
}

/* example/example_web.f90:346-391 */
void o_web_psol(const int *neq, const double *t, const double *c, const double *cdot,
                const double *savr, double *wk, const double *cj, const double *wght,
                double *rwp, int *iwp, double *b, const double *epslin, int *ierr,
                double *rpar, int *ipar)
{
    (void)t; (void)c; (void)cdot; (void)savr; (void)wght; (void)epslin; (void)rpar;
    int jpre = ipar[0];
    *ierr = 0;
    double hl0 = 1.0 / (*cj);
    if (jpre == 2 || jpre == 3) o_web_gauss_seidel(*neq, hl0, b, wk);
    if (jpre != 2) {
        if (ipar[1] == 0) o_psol_rbdpre(b, rwp, iwp);
        else o_psol_rbgpre(b, rwp, iwp); /* jbg = 1, example_web.f90:385-386 */
    }
    if (jpre == 4) o_web_gauss_seidel(*neq, hl0, b, wk);
}

/* example/example_web.f90:577-598 */
double o_web_c1_average(const double *c)
{
    double total = 0.0;
    for (int jy = 1; jy <= w_my; ++jy) {
        long iyoff = (long)w_mxns * (jy - 1);
        for (int jx = 1; jx <= w_mx; ++jx) {
            long ioff = iyoff + (long)w_ns * (jx - 1);
            total = total + c[ioff];
        }
    }
    return total / ((double)w_mx * w_my);
}

/* example/example_web.f90:600-617: root function average(c1) - 20 */
void o_web_rt(const int *neq, const double *t, const double *c, const double *cdot, const int *nrt,
              double *rval, double *rpar, int *ipar)
{
    (void)neq; (void)t; (void)cdot; (void)nrt; (void)rpar; (void)ipar;
    rval[0] = o_web_c1_average(c) - 20.0;
}

/* ------------------------------------------------------------------ */
/* heat: example/example_heat.f90                                      */
/* ------------------------------------------------------------------ */
static int h_m, h_neq;
static double h_dx, h_coeff, h_cj;

/* example/example_heat.f90:174-182 */
void o_heat_setpar(int m)
{
    h_m = m;
    h_neq = (m + 2) * (m + 2);
    h_dx = 1.0 / (m + 1);
    h_coeff = 1.0 / (h_dx * h_dx);
    h_cj = 0.0;
}

/* example/example_heat.f90:19-51 */
void o_heat_uinit(double *u, double *udot)
{
    int m = h_m;
    for (int k = 0; k <= m + 1; ++k) {
        double yk = k * h_dx;
        long ioff = (long)(m + 2) * k;
        for (int j = 0; j <= m + 1; ++j) {
            double xj = j * h_dx;
            long i = ioff + j; /* 0-based */
            u[i] = 16 * xj * (1.0 - xj) * yk * (1.0 - yk);
        }
    }
    for (long i = 0; i < h_neq; ++i) udot[i] = 0.0;
}

/* example/example_heat.f90:53-89 */
void o_heat_res(const double *t, const double *u, const double *udot, const double *cj,
                double *delta, int *ires, double *rpar, int *ipar)
{
    (void)t; (void)cj; (void)ires; (void)rpar; (void)ipar;
    const int m = h_m, m2 = h_m + 2;
    const long neq = h_neq;
    const double coeff = h_coeff;
    for (long i = 0; i < neq; ++i) delta[i] = u[i];
    for (int k = 1; k <= m; ++k) {
        long ioff = (long)m2 * k;
        for (int j = 1; j <= m; ++j) {
            long i = ioff + j; /* 0-based */
            double temx = u[i - 1] + u[i + 1];
            double temy = u[i - m2] + u[i + m2];
            delta[i] = udot[i] - (temx + temy - 4 * u[i]) * coeff;
        }
    }
}

/* example/example_heat.f90:91-109: roots at max(u) = 0.1 and 0.01 */
void o_heat_rt(const int *neq, const double *t, const double *u, const double *udot, const int *nrt,
               double *rval, double *rpar, int *ipar)
{
    (void)t; (void)udot; (void)nrt; (void)rpar; (void)ipar;
    double umax = u[0];
    for (long i = 1; i < *neq; ++i) umax = std::max(umax, u[i]);
    rval[0] = umax - 0.1;
    rval[1] = umax - 0.01;
}

/* rbdpre pairing for the heat problem (ns = nsd = 1).  NOT shipped by the
 * reference (it pairs heat with the banded preconditioner); BASELINE.json
 * config 3 asks for it, and SURVEY.md 8(d) defines it: R = -4*coeff*u on
 * interior points and R = (cj-1)*u on boundary points, so that every 1x1
 * block cj - dR/du equals the true Jacobian diagonal (cj+4*coeff, resp. 1). */
static void heat_rblock(const double *t, const int *jx, const int *jy, const double *uxy,
                        double *rxy)
{
    (void)t;
    int j = *jx - 1, k = *jy - 1;
    bool interior = (j >= 1 && j <= h_m && k >= 1 && k <= h_m);
    if (interior)
        rxy[0] = -4 * h_coeff * uxy[0];
    else
        rxy[0] = (h_cj - 1.0) * uxy[0];
}

void o_heat_jac(o_res_t res, int *ires, const int *neq, const double *t, double *u,
                double *udot, const double *rewt, double *savr, double *wk,
                const double *h, const double *cj, double *rwp, int *iwp,
                int *ierr, double *rpar, int *ipar)
{
    (void)res; (void)ires; (void)udot; (void)savr; (void)h; (void)rpar; (void)ipar;
    h_cj = *cj;
    /* r0 := R(u), held in wk(1:neq); r1 scratch is a local */
    const int m2 = h_m + 2;
    for (int jy = 1; jy <= m2; ++jy)
        for (int jx = 1; jx <= m2; ++jx) {
            long i = (long)m2 * (jy - 1) + (jx - 1);
            heat_rblock(t, &jx, &jy, &u[i], &wk[i]);
        }
    (void)neq;
    double r1[1];
    o_jac_rbdpre(*t, u, wk, heat_rblock, r1, rewt, *cj, rwp, iwp, ierr);
}

void o_heat_psol(const int *neq, const double *t, const double *u, const double *udot,
                 const double *savr, double *wk, const double *cj, const double *wght,
                 double *rwp, int *iwp, double *b, const double *epslin, int *ierr,
                 double *rpar, int *ipar)
{
    (void)neq; (void)t; (void)u; (void)udot; (void)savr; (void)wk; (void)cj; (void)wght;
    (void)epslin; (void)rpar; (void)ipar;
    *ierr = 0;
    o_psol_rbdpre(b, rwp, iwp);
}

} /* extern "C" */
