// vecapi.cu -- the whole-vector statements of DASKR's host logic as C-ABI calls on device arrays.
//
// The north star keeps the BDF / Newton control (ddstp, dnedk, dnsk, DDASIC ...) as the reference's own Fortran, calling
// a thin C layer for every statement that touches a length-NEQ array.  dkb_daskr is the C++ stand-in for that host logic
// (driver.cu); the entry points here are the same kernels exposed one statement (or one fused group of statements) at a
// time, so that a Fortran host -- fortran/daskr_b200.f90 has the bind(C) interfaces -- can replace its array statements
// call by call.  Every routine cites the reference lines it stands for.  Arrays are device pointers; scalars and the
// small coefficient arrays (psi, gama, beta) are host data, as in the reference.  Return: 0, or a negative DKB_E_* code.
#include "dkb_internal.h"
#include <limits.h>
#include <math.h>

#define VEC_ENTER()                                                                         \
    DkbCtx *c = g_ctx;                                                                      \
    if (!c) return dkb_fail(DKB_E_STATE, "dkb_init not called");                            \
    if (n < 0) return dkb_fail(DKB_E_ARG, "negative vector length")
#define VEC_LEAVE()                                                                         \
    do {                                                                                    \
        cudaError_t e__ = cudaGetLastError();                                               \
        if (e__ != cudaSuccess) return dkb_fail(DKB_E_CUDA, "%s: %s", __func__, cudaGetErrorString(e__)); \
        return 0;                                                                           \
    } while (0)

struct FMaskVt {
    const int *id;
    const double *wt;
    double *vt;
    __device__ void operator()(i64 i) const
    {
        const int m = id[i] > 0 ? id[i] : 0;
        vt[i] = (double)m * wt[i];
    }
};
template <class F>
__global__ void __launch_bounds__(256) vecapi_map_kernel(F f, i64 n)
{
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}

extern "C" {

// dcopy (extern/dblas.f:48-95) as DASKR uses it: savr = delta (dnsk:2979), phi(:,1) = y (daskr.f:1918), r = x (dslvk:3121) ...
int dkb_vcopy(int64_t n, const double *x, double *y)
{
    VEC_ENTER();
    v_copy(n, x, y);
    VEC_LEAVE();
}

// dscal (extern/dblas.f:97-136): x = a*x.  delta = confac*delta (dnsk:2975), phi(:,j) = beta(j)*phi(:,j) (ddstp:335),
// ewt = rsqrtn*ewt (dslvk:3120), v(:,ll+1) = v(:,ll+1)/snormw (dspigm:3358) ...
int dkb_vscale(int64_t n, double a, double *x)
{
    VEC_ENTER();
    v_scale(n, a, x);
    VEC_LEAVE();
}

// y = a*x (two-array form: phi(:,2) = h*yprime, daskr.f:1919)
int dkb_vscale_to(int64_t n, double a, const double *x, double *y)
{
    VEC_ENTER();
    v_scale_to(n, a, x, y);
    VEC_LEAVE();
}

// daxpy (extern/dblas.f:1-46): y = y + a*x.  x = x + z (dslvk:3143), vnew = vnew - hes*v(:,i) (dorth:3566) ...
int dkb_vaxpy(int64_t n, double a, const double *x, double *y)
{
    VEC_ENTER();
    v_axpy(n, a, x, y);
    VEC_LEAVE();
}

// z = a*x + b*y: e = e - delta (dnedk:2857 with a = 1, b = -1), delta = phi(:,kp1) + e (ddstp:367), dl = s*dl + c*v (dspigm:3340) ...
int dkb_lincomb(int64_t n, double a, const double *x, double b, const double *y, double *z)
{
    VEC_ENTER();
    v_lin2(n, a, x, b, y, z);
    VEC_LEAVE();
}

// ddot (extern/dblas.f:138-185), all-reduced over the ranks; deterministic for a given n
int dkb_ddot(int64_t n, const double *x, const double *y, double *result)
{
    VEC_ENTER();
    r_dot(n, x, y, 42);
    scal_fetch(42, 1, result);
    VEC_LEAVE();
}

// ddawts + dinvwt (src/daskr_new.f90:715-780): wt = 1/(rtol*|y| + atol).  iwt = info(2): 0 -> rtol[0], atol[0] are host
// scalars; 1 -> rtol, atol are DEVICE vectors of length n.  *ier = index (1-based, local) of the first non-positive
// weight, 0 if none (dinvwt's error return).  With vt != NULL also vt = max(id, 0)*wt (src/daskr.f:1755, 2106-2109).
int dkb_ewt_set_invert(int64_t n, int iwt, const double *rtol, const double *atol, const double *y, double *wt, double *vt,
                       const int *id, int *ier)
{
    VEC_ENTER();
    if (vt != nullptr && vt != wt && id == nullptr) return dkb_fail(DKB_E_ARG, "dkb_ewt_set_invert: vt needs the ID array");
    c->h_iflag[0] = INT_MAX;
    CUDA_TRY(cudaMemcpyAsync(c->d_iflag, c->h_iflag, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    v_ewt(n, iwt, iwt ? 0.0 : rtol[0], iwt ? 0.0 : atol[0], iwt ? rtol : nullptr, iwt ? atol : nullptr, y, wt, vt, id, c->d_iflag);
    allreduce_min_int(c->d_iflag, 1);   // several ranks: every rank must see the failure (the index is then rank-local at best)
    fetch_words(c->d_iflag, c->h_iflag, sizeof(int));
    *ier = (c->h_iflag[0] == INT_MAX) ? 0 : c->h_iflag[0];
    VEC_LEAVE();
}

// vt = max(id, 0)*wt (src/daskr.f:1755, 1881, 2106-2109; info(16) = 1)
int dkb_mask_vt(int64_t n, const int *id, const double *wt, double *vt)
{
    VEC_ENTER();
    if (n > 0) {
        const int grid = (int)std::min<i64>((n + 255) / 256, (i64)c->nsm * 8);
        vecapi_map_kernel<<<grid, 256, 0, c->stream>>>(FMaskVt{id, wt, vt}, (i64)n);
        COUNT_LAUNCH();
    }
    VEC_LEAVE();
}

// ddatrp (src/daskr_new.f90:782-826): yout, ypout = the interpolating polynomial and its derivative at tout.
// phi = base of the PHI array (column j at phi + j*ldphi, j = 0..kold), psi = host psi(1:kold+1).  One fused pass.
int dkb_interp(int64_t n, double t, double tout, int kold, const double *phi, int64_t ldphi, const double *psi, double *yout,
               double *ypout)
{
    VEC_ENTER();
    if (kold < 0 || kold > 5) return dkb_fail(DKB_E_ARG, "dkb_interp: kold out of range");
    double cc[6] = {1, 0, 0, 0, 0, 0}, dd[6] = {0, 0, 0, 0, 0, 0};
    const double dt = tout - t;
    double cv = 1.0, dv = 0.0, gama = dt / psi[0];
    for (int j = 2; j <= kold + 1; ++j) {   // :812-824
        dv = dv * gama + cv / psi[j - 2];
        cv = cv * gama;
        gama = (dt + psi[j - 2]) / psi[j - 1];
        cc[j - 1] = cv;
        dd[j - 1] = dv;
    }
    double *p[6];
    for (int j = 0; j < 6; ++j) p[j] = const_cast<double *>(phi) + (j <= kold ? j : 0) * ldphi;
    v_predict(n, kold + 1, p, cc, dd, yout, ypout);
    VEC_LEAVE();
}

// predictor (dnedk:2808-2813): y = sum_{j=1..kp1} phi(:,j), yprime = sum_{j=2..kp1} gama(j)*phi(:,j); gama = host gama(1:kp1)
int dkb_predict(int64_t n, int kp1, const double *phi, int64_t ldphi, const double *gama, double *y, double *yprime)
{
    VEC_ENTER();
    if (kp1 < 1 || kp1 > 6) return dkb_fail(DKB_E_ARG, "dkb_predict: kp1 out of range");
    double cc[6], dd[6];
    double *p[6];
    for (int j = 0; j < 6; ++j) {
        cc[j] = 1.0;
        dd[j] = (j >= 1 && j < kp1) ? gama[j] : 0.0;
        p[j] = const_cast<double *>(phi) + (j < kp1 ? j : 0) * ldphi;
    }
    v_predict(n, kp1, p, cc, dd, y, yprime);
    VEC_LEAVE();
}

// Newton update (dnsk:2988-2990): y = y - delta, e = e - delta, yprime = yprime - cj*delta
int dkb_newton_update(int64_t n, const double *delta, double cj, double *y, double *e, double *yprime)
{
    VEC_ENTER();
    v_newton_update(n, delta, cj, y, e, yprime);
    VEC_LEAVE();
}

// phi update after an accepted step (ddstp:458-466): if (write_kp2) phi(:,kp2) = e; phi(:,kp1) += e;
// phi(:,j) += phi(:,j+1) for j = kp1-1 .. 1
int dkb_phi_update(int64_t n, int kp1, int write_kp2, double *phi, int64_t ldphi, const double *e)
{
    VEC_ENTER();
    if (kp1 < 1 || kp1 > 6 || (write_kp2 && kp1 > 5)) return dkb_fail(DKB_E_ARG, "dkb_phi_update: kp1 out of range");
    double *p[7];
    for (int j = 0; j < 7; ++j) p[j] = phi + (j <= kp1 && j < 6 ? j : 0) * ldphi;
    v_phi_update(n, kp1, write_kp2, p, e);
    VEC_LEAVE();
}

// ddstp's three error-estimate norms (src/daskr_new.f90:360-378) in one pass:
// out[0] = ||e||, out[1] = ||phi(:,kp1) + e||, out[2] = ||phi(:,k) + phi(:,kp1) + e|| in the rwt-weighted RMS norm; nn = 1..3
int dkb_error_norms(int64_t n, int nn, const double *e, const double *rwt, const double *phi_kp1, const double *phi_k, double *out)
{
    VEC_ENTER();
    if (nn < 1 || nn > 3) return dkb_fail(DKB_E_ARG, "dkb_error_norms: nn must be 1..3");
    const i64 save = c->neq_g;
    if (c->nranks <= 1) c->neq_g = n;
    const bool ok = r_wrms3(n, nn, e, rwt, nn > 1 ? phi_kp1 : e, nn > 2 ? phi_k : e, out);
    c->neq_g = save;
    if (!ok) return dkb_fail(DKB_E_ARG, "dkb_error_norms: over/underflow in the plain sums; take the three ddwnrm calls instead");
    VEC_LEAVE();
}

// Device addresses of the workspace vectors dkb_daskr / dkb_alloc_workspace own, for hosts that run DASKR's own driver:
// which = 0 y, 1 yprime, 2 delta, 3 savr, 4 e, 5 wt, 6 vt, 7 phi(:,1) (ld = distance between phi columns), 8 v(:,1)
// (ld = distance between basis vectors), 9 wp, 10 iwp.  Segments of rwork in the reference (src/daskr.f:1424-1427).
int dkb_workspace_get(int which, void **ptr, int64_t *ld)
{
    DkbCtx *c = g_ctx;
    if (!c || !c->arena) return dkb_fail(DKB_E_STATE, "no workspace (dkb_daskr initial call or dkb_alloc_workspace first)");
    int64_t l = 0;
    void *p = nullptr;
    switch (which) {
    case 0: p = c->y; break;
    case 1: p = c->yp; break;
    case 2: p = c->delta; break;
    case 3: p = c->savr; break;
    case 4: p = c->e; break;
    case 5: p = c->wt; break;
    case 6: p = c->vt; break;
    case 7: p = c->phi[0]; l = c->phi[1] - c->phi[0]; break;
    case 8: p = c->v[0]; l = (c->maxl >= 1) ? c->v[1] - c->v[0] : 0; break;
    case 9: p = c->wp; break;
    case 10: p = c->iwp; break;
    default: return dkb_fail(DKB_E_ARG, "dkb_workspace_get: which = %d out of range", which);
    }
    *ptr = p;
    if (ld) *ld = l;
    return 0;
}

} // extern "C"
