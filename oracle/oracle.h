/*
 * oracle/oracle.h -- CPU restatement of the daskr Krylov hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it, and only as the checker
 * or the timed CPU baseline.  The product (daskr_b200/) never links it.
 *
 * PARITY PINNED IN PART.  The reference (HugoMVale/daskr, Fortran) cannot be compiled in this image (no
 * Fortran front end), so there is no output of the reference itself to compare with.  What the reference's
 * OWN tests pin (tests/test_oracle_krdem.py): test/test_krdem1.f90 and test/test_krdem2.f90, run here through
 * this restatement's DASKR with the Krylov option, meet the reference tests' acceptance values (solution error
 * < 1e3*atol against the analytic solution; roots at 2.47 / 2.5 / 2.53 to 1e-3 and at 81.17 + k*81.42 to 1.0).
 * That covers the host logic (DASKR, ddstp, dnedk, dnsk), dslvk/dspigm/datv/dorth/dheqr/dhels on 1x1 and 2x2
 * systems and DRCHEK/DROOTS.  PARITY UNPINNED for the rest -- the reaction-based preconditioners (rbdpre,
 * rbgpre), LINPACK, the food-web / heat routines and the Gauss-Seidel factor: the reference's tests never
 * exercise them and ship no expected outputs.  Those parts follow the reference source
 * line by line (citations are into /root/reference) and are cross-checked against independent facts only:
 * LAPACK pivots (scipy dgetrf), scipy's BLAS, numpy re-derivations of the food-web equations, of the analytic
 * block Jacobian and of the Gauss-Seidel iteration, a nonsymmetric linear system solved by dslvk, the analytic
 * heat-equation decay rate and self-consistency between preconditioner variants (tests/test_oracle_cpu.py).
 *
 * Arithmetic contract: the reference builds with gfortran for baseline x86-64,
 * i.e. separate multiply and add (no FMA contraction).  Build this with
 * -ffp-contract=off (see oracle/Makefile).
 *
 * Calling convention: Fortran-77 style, every argument by reference,
 * INTEGER = int32, REAL(rk) = double, so the signatures read like the
 * reference's (src/daskr_new.f90:11-71).
 */
#ifndef DASKR_ORACLE_H
#define DASKR_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* callback contracts, src/daskr_new.f90:11-21, 35-53, 55-71 */
typedef void (*o_res_t)(const double *t, const double *y, const double *ydot,
                        const double *cj, double *delta, int *ires,
                        double *rpar, int *ipar);
typedef void (*o_jack_t)(o_res_t res, int *ires, const int *neq, const double *t,
                         double *y, double *ydot, const double *rewt, double *savr,
                         double *wk, const double *h, const double *cj,
                         double *rwp, int *iwp, int *ierr, double *rpar, int *ipar);
typedef void (*o_psol_t)(const int *neq, const double *t, const double *y,
                         const double *ydot, const double *savr, double *wk,
                         const double *cj, const double *wght, double *rwp,
                         int *iwp, double *b, const double *epslin, int *ierr,
                         double *rpar, int *ipar);
/* rbdpre user hook, src/daskr_rbdpre.f90:176-189 */
typedef void (*o_rblock_t)(const double *t, const int *jx, const int *jy,
                           const double *uxy, double *rxy);

/* ---- BLAS-1, extern/dblas.f (stride-1 only on this path) ---- */
double o_ddot(long n, const double *dx, const double *dy);
double o_dnrm2(long n, const double *dx);
void o_daxpy(long n, double da, const double *dx, double *dy);
void o_dscal(long n, double da, double *dx);
void o_dcopy(long n, const double *sx, double *sy);
int o_idamax(int n, const double *dx);

/* ---- LINPACK, src/linpack.f90:325-431, 433-520 ---- */
void o_dgefa(double *a, int lda, int n, int *ipvt, int *info);
void o_dgesl(const double *a, int lda, int n, const int *ipvt, double *b, int job);
/* batched convenience wrappers in the bd/ipbd layout of daskr_rbdpre.f90:237-249 */
void o_dgefa_batched(double *bd, int *ipbd, int ns, long nblk, int *info);
void o_dgesl_batched(const double *bd, const int *ipbd, int ns, long nblk, double *b);

/* ---- norms / weights, src/daskr_new.f90:715-868 ---- */
double o_ddwnrm(long neq, const double *v, const double *rwt);
void o_ddawts(long neq, int iwt, const double *rtol, const double *atol,
              const double *y, double *wt);
void o_dinvwt(long neq, double *wt, int *ierr);
void o_ddatrp(double t, double tout, double *yout, double *ydotout, long neq,
              int kold, const double *phi, const double *psi);

/* ---- Krylov solver, src/daskr_new.f90:3036-3776 ---- */
void o_dheqr(double *a, int lda, int n, double *q, int *info, int ijob);
void o_dhels(const double *a, int lda, int n, const double *q, double *b);
void o_dorth(double *vnew, const double *v, double *hes, long n, int ll,
             int ldhes, int kmp, double *snormw);
void o_datv(long neq, const double *y, double t, const double *ydot,
            const double *savr, const double *v, const double *wght,
            double *ydottemp, o_res_t res, int *ires, o_psol_t psol, double *z,
            double *vtemp, double *rwp, int *iwp, double cj, double epslin,
            int *ierr, int *nres, int *npsol, double *rpar, int *ipar);
void o_dspigm(long neq, double t, const double *y, const double *ydot,
              const double *savr, double *r, const double *wght, int maxl,
              int kmp, double epslin, double cj, o_res_t res, int *ires,
              int *nres, o_psol_t psol, int *npsol, double *z, double *v,
              double *hes, double *q, int *lgmr, double *rwp, int *iwp,
              double *wk, double *dl, double *rhok, int *iflag, int irst,
              int nrsts, double *rpar, int *ipar);
void o_dslvk(long neq, const double *y, double t, const double *ydot,
             const double *savr, double *x, double *ewt, double *rwm, int *iwm,
             o_res_t res, int *ires, o_psol_t psol, int *iersl, double cj,
             double epslin, double sqrtn, double rsqrtn, double *rhok,
             double *rpar, int *ipar);

/* ---- reaction-based block-diagonal preconditioner, src/daskr_rbdpre.f90 ---- */
void o_setup_rbdpre(int mx, int my, int ns, int nsd, int lid, int *iwork);
/* src/daskr_rbgpre.f90:107-304 (block grouping) */
void o_setup_rbgpre(int mx, int my, int ns, int nsd, int nxg, int nyg, int lid, int *iwork);
void o_jac_rbgpre(double t, double *u, const double *r0, o_rblock_t rblock, double *r1,
                  const double *rewt, double cj, double *bd, int *ipbd, int *ierr);
void o_psol_rbgpre(double *b, const double *bd, const int *ipbd);
void o_jac_rbdpre(double t, double *u, const double *r0, o_rblock_t rblock,
                  double *r1, const double *rewt, double cj, double *bd,
                  int *ipbd, int *ierr);
void o_psol_rbdpre(double *b, const double *bd, const int *ipbd);

/* ---- the DASKR driver (Krylov option only), src/daskr.f:1-2493 ---- */
void o_daskr(o_res_t res, const int *neq, double *t, double *y, double *yprime,
             const double *tout, int *info, double *rtol, double *atol,
             int *idid, double *rwork, const long *lrw, int *iwork,
             const long *liw, double *rpar, int *ipar, o_jack_t jac,
             o_psol_t psol, void *rt, const int *nrt, int *jroot);

/* ---- food-web problem, example/example_web.f90 ---- */
void o_web_setpar(int np, int mx, int my);
void o_web_window(int my_window, int jy_off); /* test sampling aid, see rbdpre_problems.cpp */
void o_web_cinit(double *c, double *cdot, double pred_ic, double *rpar);
void o_web_res(const double *t, const double *c, const double *cdot,
               const double *cj, double *delta, int *ires, double *rpar, int *ipar);
void o_web_f(double t, const double *c, double *cdot, double *rpar);
void o_web_rates(const double *t, const int *jx, const int *jy, const double *c,
                 double *rxy);
void o_web_jac(o_res_t res, int *ires, const int *neq, const double *t, double *c,
               double *cdot, const double *rewt, double *savr, double *wk,
               const double *h, const double *cj, double *rwp, int *iwp,
               int *ierr, double *rpar, int *ipar);
void o_web_psol(const int *neq, const double *t, const double *c,
                const double *cdot, const double *savr, double *wk,
                const double *cj, const double *wght, double *rwp, int *iwp,
                double *b, const double *epslin, int *ierr, double *rpar, int *ipar);
void o_web_gauss_seidel(long n, double hl0, double *z, double *x);
double o_web_c1_average(const double *c);
void o_web_rt(const int *neq, const double *t, const double *c, const double *cdot, const int *nrt,
              double *rval, double *rpar, int *ipar);   /* example_web.f90:600-617 */

/* ---- heat problem, example/example_heat.f90 (+ rbdpre ns=1 pairing) ---- */
void o_heat_setpar(int m);
void o_heat_uinit(double *u, double *udot);
void o_heat_res(const double *t, const double *u, const double *udot,
                const double *cj, double *delta, int *ires, double *rpar, int *ipar);
void o_heat_rt(const int *neq, const double *t, const double *u, const double *udot, const int *nrt,
               double *rval, double *rpar, int *ipar);  /* example_heat.f90:91-109 */
void o_heat_jac(o_res_t res, int *ires, const int *neq, const double *t, double *u,
                double *udot, const double *rewt, double *savr, double *wk,
                const double *h, const double *cj, double *rwp, int *iwp,
                int *ierr, double *rpar, int *ipar);
void o_heat_psol(const int *neq, const double *t, const double *u,
                 const double *udot, const double *savr, double *wk,
                 const double *cj, const double *wght, double *rwp, int *iwp,
                 double *b, const double *epslin, int *ierr, double *rpar, int *ipar);

/* ---- whole-run helpers used by tests and the CPU baseline ---- */
/* Runs the food-web example the way example_web.f90:884-954 does (Krylov,
 * no grouping, nrt=0).  Solutions at the nout output times are stored in
 * yout[nout][neq]; iwork counters (iwork(1:40)) after each output in
 * counters[nout][40]; rwork(1:60)-header in rhead[nout][60].
 * max_steps > 0 stops after that many BDF steps (benchmark budget).  Returns idid. */
int o_run_web(int np, int mx, int my, int jpre, double rtol, double atol,
              int nout, const double *touts, double *yout, int *counters,
              double *rhead, long max_steps, double *tfinal);
/* same with block grouping (jbg = 1, nxg x nyg groups; example_web.f90:873-877 uses 5 x 5) */
int o_run_web_bg(int nxg, int nyg, int np, int mx, int my, int jpre, double rtol, double atol,
              int nout, const double *touts, double *yout, int *counters,
              double *rhead, long max_steps, double *tfinal);
void o_set_steplog(double *buf, long cap);   /* (wall s, NLI, NST) per step, max_steps mode */
long o_steplog_count(void);
int o_run_heat(int m, double rtol, double atol, int nout, const double *touts,
               double *yout, int *counters, double *rhead, long max_steps,
               double *tfinal);

#ifdef __cplusplus
}
#endif
#endif
