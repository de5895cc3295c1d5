// dkb_internal.h -- shared declarations of the daskr_b200 CUDA library.
//
// One process-wide context (the reference keeps its state in SAVE / module
// variables and is non-reentrant; we mirror that with a single
// context object bound to one GPU and one stream).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
#include "../../include/daskr_b200.h"

typedef int64_t i64;

// ---------------------------------------------------------------- limits
#define DKB_MAXL_MAX 16          // Krylov basis size the device scalar slots allow
#define DKB_NS_MAX 32            // one (sub-)warp per ns x ns block
#define DKB_RED_BLOCK 512        // threads per block in reduction kernels
#define DKB_RED_MAXGRID 1184     // 148 SMs x 8 resident blocks
#define DKB_NSLOT 64             // device scalar slots
#define DKB_GS_OV_BELOW 32        // gs.cu: rows of the slab below / above that a rank sweeps redundantly (see there)
#define DKB_GS_OV_ABOVE 8

enum { DKB_PROB_NONE = 0, DKB_PROB_WEB = 1, DKB_PROB_HEAT = 2 };

// ---------------------------------------------------------------- problems
struct WebPar {                  // module web_par, example/example_web.f90:5-60
    int np, ns, mx, my;          // my = GLOBAL rows
    double aa, ee, gg, bb, dprey, dpred, alpha, beta, ax, ay, dx, dy;
    double dyn;                  // dy/ay: row spacing in the normalised coordinate the rate factor and IC use
    double acoef[DKB_NS_MAX * DKB_NS_MAX];   // column-major (k,i) at [i*ns+k]
    double bcoef[DKB_NS_MAX], cox[DKB_NS_MAX], coy[DKB_NS_MAX];
};
struct HeatPar {                 // example/example_heat.f90:174-182
    int m;
    double dx, coeff;
};

// ---------------------------------------------------------------- profiling
struct ProfSlot {
    char name[48];
    double ms;
    i64 launches;
};
enum { PROF_DATV = 0, PROF_MGS, PROF_PSOL0, PROF_JAC, PROF_RES, PROF_VEC, PROF_NORM, PROF_HALO, PROF_TPP, PROF_RESDATV, PROF_DATV1, PROF_KIT, PROF_GS,
       PROF_NSLOTS };

// ---------------------------------------------------------------- peer memory (one node, NVLink / NVSwitch)
// Every rank owns a small mailbox in its HBM that all ranks of the node map (CUDA IPC).  Scalar all-reductions and the
// one-row halo exchange write straight into the peers' memory from inside the kernels that produce the data -- no NCCL
// call, no extra launch -- and spin on flags in their own mailbox (capi.cu: peer_*; vecops.cu: reduce_kernel's last block).
#define DKB_PEER_MAX 8
#define DKB_BOX_VAL 0          // doubles [2 parities][DKB_PEER_MAX sources][8 components]
#define DKB_BOX_ARFLAG 128     // u64 [DKB_PEER_MAX sources]: sequence number of the source's latest contribution
#define DKB_BOX_HFLAG 136      // u64 [2]: halo rows arrived from below / from above (sequence numbers)
#define DKB_BOX_WORDS 256
struct PeerComm {
    int nranks, rank;                  // nranks <= 1: no exchange
    unsigned long long seq;            // this all-reduction's sequence number (same on every rank)
    double *box[DKB_PEER_MAX];         // box[r]: rank r's mailbox as mapped in this process (box[rank]: own)
};

// ---------------------------------------------------------------- context
struct DkbCtx {
    int device = 0;
    int nsm = 148;               // multiprocessor count of the device (launch-grid sizing)
    cudaStream_t stream = nullptr;
    char err[512] = {0};
    // options
    int opt_device_io = 0;
    int opt_fused = 2;           // 0: datv through the callbacks, 1: residual + solve kernels, 2: single fused kernel
    i64 opt_max_steps = 500;
    int opt_async_out = 0;       // DKB_OPT_ASYNC_OUT: y / yprime returned through a snapshot + second stream (driver.cu)
    cudaStream_t out_stream = nullptr;
    cudaEvent_t ev_snap = nullptr, ev_out = nullptr;
    double *out_y = nullptr, *out_yp = nullptr;
    i64 out_cap = 0;
    int out_pending = 0;
    int debug_variant = 0;       // kernel-variant selector for tuning runs (dkb_debug_*)

    // communicator (NCCL, loaded lazily)
    int rank = 0, nranks = 1;
    void *comm = nullptr;
    // peer-memory path (all ranks on one node, equal slabs): mailboxes and the solver arenas of all ranks, mapped here
    int peer_ok = 0;                     // mailboxes mapped
    int peer_arena_ok = 0;               // arenas of the slab neighbours mapped (same layout on every rank)
    double *peer_box[DKB_PEER_MAX] = {nullptr};
    double *peer_arena[DKB_PEER_MAX] = {nullptr};
    double *my_box = nullptr;
    unsigned long long ar_seq = 0, halo_seq = 0;
    unsigned *halo_ticket = nullptr;

    // mesh (setup_rbdpre): mx, global my, local slab
    int rbd_set = 0;
    int mx = 0, my_g = 0, my = 0, jy0 = 0, ns = 0, nsd = 0;
    i64 npts = 0;                // local mesh points
    i64 halo = 0;                // ns*mx doubles: one mesh row
    i64 halo_pad = 0;            // room on both sides of every solver vector: one mesh row on one GPU, DKB_GS_OV_BELOW rows on several

    int problem = DKB_PROB_NONE;
    WebPar web;
    HeatPar heat;
    double *d_sx = nullptr, *d_sy = nullptr;   // sin(4 pi x_j), sin(4 pi y_k) tables (host libm)
    double heat_cj = 0.0;

    // solver vectors (device), allocated at the initial dkb_daskr call
    i64 neq = 0;                 // local unknowns
    i64 neq_g = 0;               // global unknowns (norm denominators)
    int maxl = 0, kmp = 0;
    double *arena = nullptr;
    i64 arena_doubles = 0;
    double *y = nullptr, *yp = nullptr;        // internal y / yprime (with halos)
    double *delta = nullptr, *savr = nullptr, *e = nullptr, *wt = nullptr, *vt = nullptr;
    double *phi[6] = {nullptr};
    double *v[DKB_MAXL_MAX + 1] = {nullptr};   // Krylov basis v(:,1..maxl+1); v[maxl] doubles as r
    double *dl = nullptr, *z = nullptr, *wk = nullptr, *wght = nullptr;
    double *yic = nullptr, *ypic = nullptr, *pwk = nullptr;   // IC scratch (alias v[0..2])
    double *d_rtol = nullptr, *d_atol = nullptr;              // vector tolerances (info(2)=1)
    int *d_id = nullptr;                                      // ID array (info(11)/(16))
    int *d_icnstr = nullptr;
    // preconditioner storage (WP / IWP of the reference, device resident)
    double *wp = nullptr;
    int *iwp = nullptr;
    i64 lenwp = 0, leniwp = 0;

    // device scalars + pinned mirror; reduction scratch
    // block grouping (setup_rbgpre, src/daskr_rbgpre.f90:107-193): group of every mesh column / GLOBAL row (1-based),
    // local point index of each group's representative point (-1: on another rank)
    int rbg_set = 0, ngx = 0, ngy = 0;
    int *d_jigx = nullptr, *d_jigy = nullptr;
    i64 *d_rep = nullptr;
    unsigned *gs_ticket = nullptr, *gs_done = nullptr;   // gs.cu: persistent-sweep scheduler state
    int *gs_first = nullptr;
    unsigned gs_epoch = 0;
    int gs_ntx = 0, gs_nty = 0, gs_nt = 0, gs_total = 0;
    double *gs_z = nullptr, *gs_x = nullptr;   // gs.cu: overlapped-slab work vectors (several ranks only)
    i64 gs_cap = 0;
    size_t gss_hwords = 0;                               // gs_stream.cu
    double *gss_hbuf = nullptr;                          // strip exchange buffer (all words GSS_EMPTY between applications)
    unsigned *gss_ticket = nullptr;
    double *d_scal = nullptr;    // DKB_NSLOT doubles
    double *h_scal = nullptr;    // pinned
    double *d_part = nullptr;    // DKB_RED_MAXGRID * 8 partials
    unsigned *d_ticket = nullptr;
    int *d_iflag = nullptr;      // 4 ints: error / min-index flags
    int *h_iflag = nullptr;      // pinned

    // SAVEd DASKR state (src/daskr.f:1429)
    i64 s_lid = 0;
    int s_lenid = 0, s_nonneg = 0, s_ncphi = 0, s_icnflg = 0;

    // profiling
    int prof_on = 0;
    ProfSlot prof[PROF_NSLOTS];
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_slot;    // slot of pair i (events 2i, 2i+1)
    i64 launches = 0;
    i64 reorth_entered = 0, reorth_corrections = 0;   // dorth:3575-3588 branch counters (tests)
    i64 ll_hist[DKB_MAXL_MAX + 1] = {0};   // Krylov iterations taken at basis index ll (bench.py: whole-iteration roofline)
};

extern DkbCtx *g_ctx;

// ---------------------------------------------------------------- errors
int dkb_fail(int code, const char *fmt, ...);
#define CUDA_TRY(x)                                                                       \
    do {                                                                                  \
        cudaError_t e__ = (x);                                                            \
        if (e__ != cudaSuccess)                                                           \
            return dkb_fail(DKB_E_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #x,          \
                            cudaGetErrorString(e__));                                     \
    } while (0)
void dkb_cuda_check(const char *what);   // aborts loudly on a CUDA error (void paths)

// profiling helpers
void prof_begin(int slot);
void prof_end(int slot);
struct ProfScope {
    int slot;
    explicit ProfScope(int s) : slot(s) { prof_begin(s); }
    ~ProfScope() { prof_end(slot); }
};
#define COUNT_LAUNCH() (g_ctx->launches++)

// ---------------------------------------------------------------- vector ops (vecops.cu)
// Elementwise ("K7 lincomb family") -- all on g_ctx->stream.
void v_copy(i64 n, const double *x, double *y);
void v_zero(i64 n, double *y);
void v_scale(i64 n, double a, double *x);                                  // x = a*x
void v_scale_to(i64 n, double a, const double *x, double *y);              // y = a*x
void v_axpy(i64 n, double a, const double *x, double *y);                  // y = y + a*x
void v_lin2(i64 n, double a, const double *x, double b, const double *y, double *z); // z = a*x + b*y
void v_mul(i64 n, const double *x, const double *w, double *y);            // y = x*w
void v_sub(i64 n, const double *x, const double *y, double *z);            // z = x - y
// multi-vector: yout = p1 + sum_j c_j p_j, dout = sum_j d_j p_j  (predictor dnedk:2808-2813, ddatrp:812-824)
void v_predict(i64 n, int kp1, double *const *phi, const double *c, const double *d,
               double *yout, double *dout);
// phi update cascade, ddstp:458-466
void v_phi_update(i64 n, int kp1, int write_kp2, double *const *phi, const double *e);
// ewt: wt = 1/(rtol*|y|+atol) (ddawts+dinvwt, :715-780), vt = max(id,0)*wt (daskr.f:1755).
// Returns first non-positive weight index (1-based, local) in *ierr_local or 0.
void v_ewt(i64 n, int iwt, double rtol, double atol, const double *d_rtol, const double *d_atol,
           const double *y, double *wt, double *vt, const int *id, int *first_bad);
// datv unfused pieces (:3493-3498, 3509, 3518)
void v_datv_pre(i64 n, const double *v, double vscale, const double *wt, double wscale, const double *y,
                const double *ydot, double cj, double *vtemp, double *ydottemp, double *z);
// Newton update (dnsk:2988-2993): y-=delta, e-=delta, ydot-=cj*delta; returns wrms partials
void v_newton_update(i64 n, const double *delta, double cj, double *y, double *e, double *ydot);
// IC helpers
void v_dyypnw(i64 n, const double *y, const double *yp, double cj, double rl, const double *p,
              int icopt, const int *id, double *ynew, double *ypnew);
void v_min0(i64 n, const double *y, double *delta);                       // delta = min(y,0)
// solution assembly (dspigm:3413-3417) z = (sum_i r_i v_i)/(wscale*wt); x = x + z (or x = z)
void v_gmres_solution(i64 n, int ll, double *const *v, const double *vs, const double *r, const double *wt,
                      double wscale, double *x, int first);
// restart residual (dspigm:3381-3401): dl = (v1; dl = s*dl + c*v_{i+1} ...) * scale
void v_gmres_dl(i64 n, int maxl, double *const *v, const double *vs, const double *q, double snormw, double scale,
                double *dl);
// the same with a normalisation factor on x / y: the Krylov basis is kept unnormalised (krylov.cu)
void v_axpy_s(i64 n, double a, const double *x, double s, double *y);                                 // y = y + a*(s*x)
void v_lin2_s(i64 n, double a, const double *x, double b, const double *y, double s, double *z);      // z = a*x + b*(s*y)
void r_dot_s(i64 n, const double *x, double sx, const double *y, int slot);                           // (sx*x).y

// Reductions: results land in g_ctx->d_scal[slot...] (allreduced across ranks), host reads
// them with scal_fetch().  All deterministic for a fixed n / grid.
void r_dot(i64 n, const double *x, const double *y, int slot);                 // ddot
void r_sum_strided(i64 npts, i64 stride, i64 off, const double *x, int slot);  // sum_i x[i*stride+off] (root functions)
void r_max(i64 n, const double *x, int slot);
bool r_wrms3(i64 n, int nn, const double *e, const double *rwt, const double *phi_kp1, const double *phi_k, double *out);   // ddstp error estimates, one pass
void r_cnstr(i64 n, const double *y, const double *ynew, const int *icnstr, int slot);   // dcnstr: [rdymx, hard flag]
void v_cnst0(i64 n, const double *y, const int *icnstr, int *d_flag);                   // dcnst0: atomicMin(first bad index + 1)                                  // max_i x[i]
void r_dot2(i64 n, const double *vnew, const double *v1, double s1, int slot);            // [vnew.vnew, (s1*v1).vnew]
void r_axpy_dot(i64 n, double *vnew, const double *vi, double si, int hslot, const double *vnext, double sn,
                int slot);   // vnew -= h*(si*vi) ; out = (sn*vnext).vnew
void r_axpy_nrm(i64 n, double *vnew, const double *vi, double si, int hslot, int slot);  // vnew -= h*(si*vi) ; out = vnew.vnew
void r_wrms_parts(i64 n, const double *v, const double *rwt, int slot);        // [sum (v*rwt)^2 , max|v*rwt|]
void scal_fetch(int slot, int count, double *out);   // stream-ordered readback through mapped memory + sync
void fetch_words(const void *dsrc, void *hdst, size_t bytes);   // same for any few words; hdst in h_scal / h_iflag
double r_wrms(i64 n, const double *v, const double *rwt);   // ddwnrm :828-868, global N
double r_nrm2(i64 n, const double *x);
void allreduce_sum(int slot, int count);
void allreduce_max(int slot, int count);
void allreduce_min_int(int *d_val, int count);
PeerComm peer_comm_next();        // capi.cu: descriptor of the next in-kernel all-reduction (nranks = 1 if the peer path is off)
int peer_map_arena();             // capi.cu: (re)map the neighbours' solver arenas after alloc_solver (collective)

// ---------------------------------------------------------------- LU (lu.cu)
int lu_dgefa_batched(double *bd, int *ipbd, int ns, i64 nblk, int *info, int *d_first_flag);
int lu_dgesl_batched(const double *bd, const int *ipbd, int ns, i64 nblk, double *b);
// psol with fused scaling: out = wscale*wt * P^-1 in  (dspigm prologue :3278-3281) + sum of squares
int lu_psol_scaled(const double *bd, const int *ipbd, int ns, i64 nblk, const double *in,
                   const double *wt, double wscale, double *out, int slot);

// thread-per-point block solve over the interleaved LU layout (psol_tpp.cu)
int tpp_psol(const double *lut, const int *pt, int ns, i64 npts, const double *in, const double *savr,
             const double *wt, double wscale, double *out, int slot);
i64 tpp_lut_doubles(int ns, i64 npts);
i64 tpp_pt_ints(int ns, i64 npts);
int tpp_from_reference(int ns, i64 npts, const double *bd, const int *ipbd, double *lut, int *pt);
int tpp_to_reference(int ns, i64 npts, const double *lut, const int *pt, double *bd, int *ipbd);

// ---------------------------------------------------------------- problems (web.cu / heat.cu)
void web_upload_tables();
void web_res_launch(const double *c, const double *cdot, double *delta);
int web_gauss_seidel(double hl0, double *z, double *x);
int gs_stream_sweep(double hl0, int my, bool bot_bnd, bool top_bnd, const double *zin, double *zout, int npanels);   // gs_stream.cu: the five sweeps in one pass
// the factor applied to z; *res = where the result is (z itself, or x when the sweep ran out of place); x: N words of scratch
int web_gauss_seidel_to(double hl0, double *z, double *x, double **res);
// block-grouped reaction factor (jbg = 1): FD blocks at the representative points (unfactored, reference layout)
void web_jac_groups_launch(const double *c, const double *rewt, double *bd);
void rbg_add_cj(double *bd, int ns, int nsd, int ngrp, double cj);
int rbg_psol(const double *bd, const int *ipbd, double *b);
void allreduce_sum_buf(double *buf, i64 n);   // gs.cu: z := GS(A_S) z, x = N words of scratch
// returns true if the speculative kernel ran (jac3.cu): a flagged zero pivot then means "run again with exact = true"
bool web_jac_launch(const double *c, const double *rewt, double cj, double *bd, int *ipbd,
                    int *d_first_flag, bool exact = false);
// fused datv: vnext = wscale*wt * P^-1 ( G(y + v/(wscale*wt), ydot + cj*v/(wscale*wt)) - savr )
int web_datv_fused(const double *v, double vscale, const double *wt, double wscale, const double *y, const double *ydot,
                   const double *savr, double cj, const double *bd, const int *ipbd, double *vnext, int jpre);
void heat_res_launch(const double *u, const double *udot, double *delta);
void heat_jac_launch(const double *u, const double *rewt, double cj, double *bd, int *ipbd,
                     int *d_first_flag);
void heat_datv_fused(const double *v, double vscale, const double *wt, double wscale, const double *y,
                     const double *ydot, const double *savr, double cj, const double *bd,
                     double *vnext);

// halo exchange of one mesh row each way (K9); no-op without a communicator
void halo_exchange(double *vec);
void halo_exchange3(double *a, double *b, double *c);
// overlap rows for the slab-wise Gauss-Seidel sweep: my top n_up doubles go to rank+1 (arriving in its from_below),
// my bottom n_down doubles to rank-1 (arriving in its from_above)
void slab_overlap_exchange(const double *own, i64 n_own, i64 n_up, i64 n_down, double *from_below, double *from_above);

// ---------------------------------------------------------------- krylov.cu
void k_dslvk(i64 neq, const double *y, double t, const double *ydot, const double *savr,
             double *x, const double *ewt, int *iwm, dkb_res_t res, int *ires, dkb_psol_t psol,
             int *iersl, double cj, double epslin, double sqrtn, double rsqrtn, double *rhok,
             double *rpar, int *ipar);
bool builtin_fused(dkb_res_t res, dkb_psol_t psol, const int *ipar);
