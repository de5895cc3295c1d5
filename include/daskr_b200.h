/*
 * daskr_b200.h -- C ABI of the B200-native Krylov path of DASKR.
 *
 * Drop-in boundary for HugoMVale/daskr's preconditioned-GMRES inner loop
 * (dslvk/dspigm/datv/dorth/ddwnrm + the reaction-based block-diagonal
 * preconditioner jac_rbdpre/psol_rbdpre with batched dgefa/dgesl).  The
 * reference has no FFI layer: its boundary is the Fortran-77 procedure
 * convention itself (every argument by reference, INTEGER = int32,
 * REAL(rk) = double, callbacks = C function pointers).  Every entry point
 * below keeps that convention so a Fortran host binds it with
 * `bind(C, name=...)` interfaces (INTEGRATION.md shows the stubs) and names
 * the reference interface it replaces.
 *
 * What changes against the reference: all length-NEQ arrays that the
 * callbacks see (y, ydot, delta, savr, wk, wght, b, rwp, iwp) are DEVICE
 * pointers, the solver's own vectors (phi, Krylov basis, LU blocks) live in
 * device memory owned by the library, and rwork/iwork shrink to their scalar
 * headers (rwork(1:60), iwork(1:40) + the ID / constraint arrays).  The
 * y/yprime arguments of dkb_daskr are host arrays (copied in on the initial
 * call, out at every return) unless DKB_OPT_DEVICE_IO is set.
 *
 * Like the reference (SAVE / module variables) the library is
 * non-reentrant: one problem per process, bound to one GPU.  All functions
 * return 0 on success or a negative DKB_E_* code (dkb_last_error() has text),
 * except the Fortran-style subroutines which report through their own
 * ires / ierr / idid arguments exactly like the routines they replace.
 */
#ifndef DASKR_B200_H
#define DASKR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DKB_VERSION 100

enum {
    DKB_OK = 0,
    DKB_E_CUDA = -1,     /* a CUDA runtime call failed */
    DKB_E_ARG = -2,      /* illegal argument */
    DKB_E_STATE = -3,    /* call out of order (no context, no setup, ...) */
    DKB_E_UNSUPPORTED = -4,
    DKB_E_NCCL = -5
};

/* options for dkb_set_option */
enum {
    DKB_OPT_DEVICE_IO = 1,  /* 1: y/yprime passed to dkb_daskr are device pointers */
    DKB_OPT_FUSED = 2,      /* built-in problems: 2 (default) single-kernel datv, 1 two kernels, 0 plain callbacks */
    DKB_OPT_MAX_STEPS = 3,  /* per-call step limit (reference: 500, src/daskr.f:2057) */
    DKB_OPT_ASYNC_OUT = 4   /* 1: successful returns hand y / yprime to the host through a device snapshot and a second
                             * stream, overlapping the PCIe transfer with the next call's step; the arrays (pinned host
                             * memory) are complete after dkb_output_wait() or after the next dkb_daskr call returns */
};

/* ---- callback contracts: src/daskr_new.f90:11-21, 35-53, 55-71 ----
 * Same argument lists as res_t / jack_t / psol_t; array arguments are device
 * pointers, scalars are host pointers, rpar/ipar are the caller's host arrays
 * passed through untouched.  Callbacks enqueue their work on dkb_stream(). */
typedef void (*dkb_res_t)(const double *t, const double *y, const double *ydot,
                          const double *cj, double *delta, int *ires,
                          double *rpar, int *ipar);
typedef void (*dkb_jack_t)(dkb_res_t res, int *ires, const int *neq, const double *t,
                           double *y, double *ydot, const double *rewt, double *savr,
                           double *wk, const double *h, const double *cj,
                           double *rwp, int *iwp, int *ierr, double *rpar, int *ipar);
typedef void (*dkb_psol_t)(const int *neq, const double *t, const double *y,
                           const double *ydot, const double *savr, double *wk,
                           const double *cj, const double *wght, double *rwp,
                           int *iwp, double *b, const double *epslin, int *ierr,
                           double *rpar, int *ipar);

/* rblock of jac_rbdpre, src/daskr_rbdpre.f90:176-189, batched over the mesh: r = R(t, u) for ALL mesh points in one call
 * (u, r device arrays of ns*mx*my doubles, species fastest, my = this rank's rows; one launch on dkb_stream()).
 * `user` is passed through untouched. */
typedef void (*dkb_rblock_t)(const double *t, const int *mx, const int *my, const int *ns, const double *u, double *r,
                             void *user);

/* root functions, src/daskr.f:908-934: rt(neq,t,y,yp,nrt,rval,rpar,ipar); y, yp device, rval host */
typedef void (*dkb_rt_t)(const int *neq, const double *t, const double *y, const double *yp,
                         const int *nrt, double *rval, double *rpar, int *ipar);

/* ---- context (replaces the reference's SAVE/module state) ---- */
int dkb_init(int device);            /* create the process-wide context on `device` */
int dkb_finalize(void);
const char *dkb_last_error(void);
int dkb_version(void);
void *dkb_stream(void);              /* cudaStream_t the library launches on */
int dkb_set_option(int opt, int64_t value);
int dkb_sync(void);
int dkb_output_wait(void);           /* DKB_OPT_ASYNC_OUT: block until the last returned y / yprime have arrived */

/* ---- multi-GPU: y-slab decomposition, NCCL halo exchange + scalar allreduce ----
 * dkb_comm_unique_id writes an ncclUniqueId (128 bytes) on rank 0; every rank
 * then calls dkb_comm_init with the same id. */
int dkb_comm_unique_id(void *id128);
int dkb_comm_init(const void *id128, int rank, int nranks);

/* ---- setup_rbdpre, src/daskr_rbdpre.f90:102-156 (DMSET2) ----
 * Same arguments.  my is the GLOBAL number of mesh rows; with a communicator
 * this rank owns rows [jy0, jy0+myloc) (dkb_slab reports them) and NEQ passed
 * to dkb_daskr is ns*mx*myloc.  iwork(27:28) (LENWP, LENIWP) are set to 0:
 * the blocks and pivots live in device memory sized in 64 bits
 * (ns*ns*mx*my overflows INTEGER at 4096^2 x 20).  If lid != 0 the ID array
 * is written to iwork(lid+1 : lid+neq) exactly as the reference does. */
void dkb_setup_rbdpre(const int *mx, const int *my, const int *ns, const int *nsd,
                      const int *lid, int *iwork);
int dkb_slab(int *jy0, int *myloc);

/* setup_rbgpre (DGSET2), src/daskr_rbgpre.f90:107-158: the block-grouped variant, nxg x nyg groups of mesh points
 * sharing one block (the block of the group's representative point).  Same arguments as the reference; replaces
 * dkb_setup_rbdpre for runs with jbg = ipar(2) = 1 of the food-web routines (example_web.f90:873-877). */
void dkb_setup_rbgpre(const int *mx, const int *my, const int *ns, const int *nsd, const int *nxg, const int *nyg,
                      const int *lid, int *iwork);

/* ---- built-in device problems (the user routines of the two examples) ---- */
/* food web, example/example_web.f90:5-60 (setpar) with np, mx, my at run time */
int dkb_web_setpar(int np, int mx, int my);
/* the same problem on [0,1] x [0,ay] (fields in y/ay, true spacing in the stencil); ay=1 is setpar */
int dkb_web_setpar_ex(int np, int mx, int my, double ay);
/* res / jac / psol of example_web.f90:190-224, 301-344, 346-391.  As in the example, ipar(1) = jpre selects the
 * preconditioner -- 1: reaction factor A_R (rbdpre / rbgpre), 2: spatial Gauss-Seidel factor A_S, 3: A_S then A_R
 * (the shipped setting), 4: A_R then A_S -- and ipar(2) = jbg selects block grouping (0: dkb_setup_rbdpre,
 * 1: dkb_setup_rbgpre; must match the setup call that was made).  psol uses wk as N words of scratch.
 * On several ranks (dkb_comm_init) res exchanges boundary rows with the neighbouring slabs IN PLACE: one mesh row is
 * written in front of and behind c, room that only the solver's own vectors have (dkb_daskr's workspace,
 * dkb_workspace_get).  Called on any other array it sets ires = -2 and explains on stderr; the same holds for
 * dkb_heat_res. */
void dkb_web_res(const double *t, const double *c, const double *cdot, const double *cj,
                 double *delta, int *ires, double *rpar, int *ipar);
void dkb_web_jac(dkb_res_t res, int *ires, const int *neq, const double *t, double *c,
                 double *cdot, const double *rewt, double *savr, double *wk,
                 const double *h, const double *cj, double *rwp, int *iwp, int *ierr,
                 double *rpar, int *ipar);
void dkb_web_psol(const int *neq, const double *t, const double *c, const double *cdot,
                  const double *savr, double *wk, const double *cj, const double *wght,
                  double *rwp, int *iwp, double *b, const double *epslin, int *ierr,
                  double *rpar, int *ipar);
/* rt, example_web.f90:600-617: rval(1) = average(c1) - 20 */
void dkb_web_rt(const int *neq, const double *t, const double *c, const double *cdot, const int *nrt,
                double *rval, double *rpar, int *ipar);
/* cinit, example_web.f90:102-148: fills HOST arrays c, cdot for this rank's slab */
int dkb_web_cinit(double *c, double *cdot, double pred_ic);
/* heat, example/example_heat.f90:53-89 with the rbdpre ns=1 pairing of SURVEY 8(d) */
int dkb_heat_setpar(int m);
void dkb_heat_res(const double *t, const double *u, const double *udot, const double *cj,
                  double *delta, int *ires, double *rpar, int *ipar);
void dkb_heat_jac(dkb_res_t res, int *ires, const int *neq, const double *t, double *u,
                  double *udot, const double *rewt, double *savr, double *wk,
                  const double *h, const double *cj, double *rwp, int *iwp, int *ierr,
                  double *rpar, int *ipar);
void dkb_heat_psol(const int *neq, const double *t, const double *u, const double *udot,
                   const double *savr, double *wk, const double *cj, const double *wght,
                   double *rwp, int *iwp, double *b, const double *epslin, int *ierr,
                   double *rpar, int *ipar);
int dkb_heat_uinit(double *u, double *udot);
/* rt, example_heat.f90:91-109: rval = max(u) - 0.1, max(u) - 0.01 */
void dkb_heat_rt(const int *neq, const double *t, const double *u, const double *udot, const int *nrt,
                 double *rval, double *rpar, int *ipar);

/* ---- DASKR, src/daskr.f:1-3 (Krylov option info(12)=1) ----
 * Same 21 arguments.  lrw/liw are 64-bit (lrw >= 60 + 3*nrt).  rt is a dkb_rt_t (or NULL with nrt = 0). */
void dkb_daskr(dkb_res_t res, const int *neq, double *t, double *y, double *yprime,
               const double *tout, int *info, double *rtol, double *atol, int *idid,
               double *rwork, const int64_t *lrw, int *iwork, const int64_t *liw,
               double *rpar, int *ipar, dkb_jack_t jac, dkb_psol_t psol, void *rt,
               const int *nrt, int *jroot);

/* ---- checkpoint-style continuation.  The reference keeps its whole state in the caller's
 * rwork / iwork (src/daskr.f:1424-1427, 1724-1735): saving those arrays is a checkpoint.  Here their length-NEQ
 * segments (PHI, WP/IWP, ID ...) live on the device; these calls move them to / from one HOST buffer.  Save
 * rwork(1:60+3*nrt), iwork and (t, idid) next to it; after dkb_init + dkb_setup_rbdpre + setpar in a new process,
 * dkb_state_import and a continuation call (info(1) = 1) take the same steps as the uninterrupted run. */
int64_t dkb_state_bytes(void);
int dkb_state_export(void *host_buf, int64_t nbytes);
int dkb_state_import(const void *host_buf, int64_t nbytes);

/* ---- the hot path as separate calls (what a Fortran dnsk/dnedk would bind) ---- */
/* The device workspace dkb_daskr allocates on its initial call (solver vectors, Krylov basis v(:,1:maxl+1), WP/IWP),
 * for hosts that keep DASKR's own driver and call dkb_dslvk / the vector-op layer directly: the part of
 * rwork(LWM:) / iwork(LIWM:) that src/daskr.f:1560-1604 carves for the Krylov option.  Call after dkb_setup_rbdpre. */
int dkb_alloc_workspace(int64_t neq, int maxl, int kmp, int64_t lenwp, int64_t leniwp);
/* Device addresses of the workspace vectors (segments of rwork in the reference, src/daskr.f:1424-1427):
 * which = 0 y, 1 yprime, 2 delta, 3 savr, 4 e, 5 wt, 6 vt, 7 phi(:,1), 8 v(:,1) (Krylov basis), 9 wp, 10 iwp;
 * *ld = distance in doubles between consecutive columns of phi / v (0 for the others). */
int dkb_workspace_get(int which, void **ptr, int64_t *ld);

/* ---- the whole-vector statements of the host logic on device arrays: what the Fortran ddstp / dnedk /
 * dnsk / DDASIC replace their array statements with, one call per statement or fused group (csrc/vecapi.cu cites the
 * reference lines of each).  n = local vector length; scalars and psi / gama are host data. ---- */
int dkb_vcopy(int64_t n, const double *x, double *y);                               /* dcopy, extern/dblas.f:48 */
int dkb_vscale(int64_t n, double a, double *x);                                     /* dscal, extern/dblas.f:97 */
int dkb_vscale_to(int64_t n, double a, const double *x, double *y);                 /* y = a*x, daskr.f:1919 */
int dkb_vaxpy(int64_t n, double a, const double *x, double *y);                     /* daxpy, extern/dblas.f:1 */
int dkb_lincomb(int64_t n, double a, const double *x, double b, const double *y, double *z);   /* z = a*x + b*y */
int dkb_ddot(int64_t n, const double *x, const double *y, double *result);          /* ddot, extern/dblas.f:138 */
/* ddawts + dinvwt, src/daskr_new.f90:715-780 (+ the vt mask of daskr.f:1755 when vt != NULL).  iwt = info(2): 0 ->
 * rtol, atol host scalars; 1 -> device vectors.  *ier = first non-positive weight (1-based) or 0. */
int dkb_ewt_set_invert(int64_t n, int iwt, const double *rtol, const double *atol, const double *y, double *wt,
                       double *vt, const int *id, int *ier);
int dkb_mask_vt(int64_t n, const int *id, const double *wt, double *vt);            /* vt = max(id,0)*wt, daskr.f:1755 */
/* ddatrp, src/daskr_new.f90:782-826: phi column j at phi + j*ldphi, psi = host psi(1:kold+1) */
int dkb_interp(int64_t n, double t, double tout, int kold, const double *phi, int64_t ldphi, const double *psi,
               double *yout, double *ypout);
/* predictor, dnedk:2808-2813: y = sum phi(:,j), yprime = sum gama(j)*phi(:,j); gama = host gama(1:kp1) */
int dkb_predict(int64_t n, int kp1, const double *phi, int64_t ldphi, const double *gama, double *y, double *yprime);
int dkb_newton_update(int64_t n, const double *delta, double cj, double *y, double *e, double *yprime);   /* dnsk:2988-2990 */
int dkb_phi_update(int64_t n, int kp1, int write_kp2, double *phi, int64_t ldphi, const double *e);       /* ddstp:458-466 */
/* ddstp:360-378: ||e||, ||phi(:,kp1)+e||, ||phi(:,k)+phi(:,kp1)+e|| (nn = 1..3 of them) in one pass */
int dkb_error_norms(int64_t n, int nn, const double *e, const double *rwt, const double *phi_kp1, const double *phi_k,
                    double *out);

/* dslvk, src/daskr_new.f90:3036-3165.  Device arrays y, ydot, savr, x, ewt of
 * length neq; counters are accumulated into iwm(12,16,20,21) like the reference. */
void dkb_dslvk(const int *neq, const double *y, const double *t, const double *ydot,
               const double *savr, double *x, double *ewt, int *iwm, dkb_res_t res,
               int *ires, dkb_psol_t psol, int *iersl, const double *cj,
               const double *epslin, const double *sqrtn, const double *rsqrtn,
               double *rhok, double *rpar, int *ipar);
/* ddwnrm, src/daskr_new.f90:828-868 (device v, rwt) */
double dkb_ddwnrm(const int *neq, const double *v, const double *rwt);
/* jac_rbdpre, src/daskr_rbdpre.f90:158-251, for any user reaction term: same arguments plus the `user` pointer handed to
 * the batched rblock; r1 is N doubles of scratch.  Writes bd / ipbd in the reference's layout (block p at bd + p*ns*ns,
 * column-major; pivots 1-based at ipbd + p*ns); *ierr = dgefa's info of the first failing block (0: none). */
void dkb_jac_rbdpre(const double *t, double *u, const double *r0, dkb_rblock_t rblock, void *user, double *r1,
                    const double *rewt, const double *cj, double *bd, int *ipbd, int *ierr);
/* psol_rbdpre on the tile-interleaved blocks (dkb_rbd_from_reference / the built-in jac routines): streaming solve */
void dkb_psol_rbdpre_tiles(double *b, const double *rwp, const int *iwp);
/* psol_rbdpre, src/daskr_rbdpre.f90:253-279, on device arrays in the reference's bd / ipbd
 * layout (block p at bd+p*ns*ns column-major, pivots 1-based at ipbd+p*ns) */
void dkb_psol_rbdpre(double *b, const double *bd, const int *ipbd);
/* The rwp / iwp arrays the jac / psol callbacks see hold the same LU factors and pivots in a
 * tile-interleaved order (32 mesh points per tile, see psol_tpp.cu) so that the per-point
 * solves stream at HBM speed.  dkb_rbd_lenwp / leniwp give their lengths (doubles / ints);
 * the two converters map between that order and the reference's bd / ipbd (device arrays). */
int64_t dkb_rbd_lenwp(void);
int64_t dkb_rbd_leniwp(void);
int dkb_rbd_to_reference(const double *rwp, const int *iwp, double *bd, int *ipbd);
int dkb_rbd_from_reference(const double *bd, const int *ipbd, double *rwp, int *iwp);

/* ---- batched LINPACK micro-API (BASELINE config 5), src/linpack.f90:325-520 ----
 * Device arrays in the bd / ipbd layout.  info_first receives the 1-based index
 * of the first block whose dgefa info != 0 (0 if none); info (optional, device,
 * nblk ints) receives every block's dgefa info. */
int dkb_dgefa_batched(double *bd, int *ipbd, int ns, int64_t nblk, int *info,
                      int64_t *info_first);
int dkb_dgesl_batched(const double *bd, const int *ipbd, int ns, int64_t nblk, double *b);

/* ---- device memory helpers for hosts without a CUDA runtime binding ---- */
int dkb_malloc(void **ptr, int64_t bytes);
int dkb_free(void *ptr);
int dkb_memcpy_h2d(void *dst, const void *src, int64_t bytes);
int dkb_memcpy_d2h(void *dst, const void *src, int64_t bytes);
int dkb_memcpy_d2d(void *dst, const void *src, int64_t bytes);
int dkb_memset(void *dst, int value, int64_t bytes);

/* ---- measurement hooks (bench.py): kernel-time accumulation by CUDA events ---- */
int dkb_prof_enable(int on);
int dkb_prof_reset(void);
/* name[i] (up to 32 entries of 48 chars), accumulated ms, launch counts */
int dkb_prof_get(char *names, double *ms, int64_t *launches, int max_entries);
int64_t dkb_launch_count(void);
/* Krylov iterations since dkb_prof_reset by basis index: hist[ll-1], ll = 1..n (dspigm's loop, src/daskr_new.f90:3304-3361) */
int dkb_krylov_stats(int64_t *hist, int n);
/* dorth's conditional reorthogonalisation (src/daskr_new.f90:3575-3588): branch entered / corrections applied so far */
int dkb_reorth_stats(int64_t *entered, int64_t *corrections);

#ifdef __cplusplus
}
#endif
#endif
