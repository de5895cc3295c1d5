// web.cu -- device kernels of the food-web problem (example/example_web.f90)
//   K4  residual           res:190-224 + f:226-269 + rates:271-299
//   K5  block Jacobian     jac:301-344 -> jac_rbdpre (src/daskr_rbdpre.f90:158-251) + dgefa
//   K3  fused datv         datv (src/daskr_new.f90:3493-3518) with res and psol_rbdpre inlined
//
// Mapping: one sub-warp group (G = 2^ceil(log2 ns) lanes) per mesh point, lane i
// = species i.  Unknown (i,jx,jy) sits at i + ns*(jx-1) + ns*mx*(jy-1), so a
// group touches ns*8 contiguous bytes per array and per neighbour; x-neighbours
// are the adjacent groups' lines (L1), y-neighbours one mesh row away (L2).
// The ns x ns interaction sum uses shuffle broadcasts of the point's
// concentrations against a shared-memory copy of acoef; the LU blocks of the
// point stream through registers (lu.cu layout) exactly once per launch.
// Mesh rows are slab-decomposed: rows outside the slab are read from the halo
// rows just before / after the local array, global edges use the reference's
// mirrored Neumann differences.
#include "dkb_internal.h"
#include <math.h>
#include <string.h>

#include "lu_device.cuh"   // group_dgefa / group_dgesl / group_load_block

struct WebDev {
    int ns, np, mx, my, jy0, my_g, has_lo, has_hi;
    double alpha, beta, dx, dy, srur;
    const double *acoef, *bcoef, *cox, *coy, *sx, *sy;   // device tables
};

static WebDev make_webdev()
{
    DkbCtx *c = g_ctx;
    WebDev w;
    w.ns = c->web.ns;
    w.np = c->web.np;
    w.mx = c->mx;
    w.my = c->my;
    w.jy0 = c->jy0;
    w.my_g = c->my_g;
    w.has_lo = c->jy0 > 0;
    w.has_hi = c->jy0 + c->my < c->my_g;
    w.alpha = c->web.alpha;
    w.beta = c->web.beta;
    w.dx = c->web.dx;
    w.dy = c->web.dy;
    w.srur = 1.4901161193847656e-08;   // sqrt(epsilon(one)) = 2^-26, daskr_rbdpre.f90:129
    const double *tab = c->d_sx;       // [sx(mx) | sy(my_g) | acoef(ns*ns) | bcoef | cox | coy]
    w.sx = tab;
    w.sy = tab + c->mx;
    w.acoef = w.sy + c->my_g;
    w.bcoef = w.acoef + (i64)w.ns * w.ns;
    w.cox = w.bcoef + w.ns;
    w.coy = w.cox + w.ns;
    return w;
}

// sin(4 pi x_j) and sin(4 pi y_k) are evaluated on the host with the same libm the CPU path
// uses, so fac = 1 + alpha*x*y + beta*sin*sin (rates:286-288) is bit-identical on both sides.
void web_upload_tables()
{
    DkbCtx *c = g_ctx;
    const WebPar &p = c->web;
    const double pi = 4 * atan(1.0);
    std::vector<double> tab;
    for (int jx = 1; jx <= c->mx; ++jx) tab.push_back(sin(4 * pi * ((jx - 1) * p.dx)));
    for (int jy = 1; jy <= c->my_g; ++jy) tab.push_back(sin(4 * pi * ((jy - 1) * p.dy)));
    for (int i = 0; i < p.ns * p.ns; ++i) tab.push_back(p.acoef[i]);
    for (int i = 0; i < p.ns; ++i) tab.push_back(p.bcoef[i]);
    for (int i = 0; i < p.ns; ++i) tab.push_back(p.cox[i]);
    for (int i = 0; i < p.ns; ++i) tab.push_back(p.coy[i]);
    if (c->d_sx) cudaFree(c->d_sx);
    cudaMalloc(&c->d_sx, tab.size() * sizeof(double));
    cudaMemcpy(c->d_sx, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice);
    dkb_cuda_check("web_upload_tables");
}

// ---- accessors for the state the residual is evaluated at ----
struct StatePlain {       // G(t, c, cdot)
    const double *c;
    __device__ __forceinline__ double operator()(i64 idx) const { return __ldg(c + idx); }
};
struct StateDatv {        // z = y + v/wght, wght = wscale*wt   (datv:3493-3498)
    const double *y, *v, *wt;
    double wscale;
    __device__ __forceinline__ double operator()(i64 idx) const
    {
        return __ldg(y + idx) + __ldg(v + idx) / (wscale * __ldg(wt + idx));
    }
};

// rates:286-297 for lane li given the point's concentrations (ci own, broadcast by shuffle)
template <int NS>
__device__ __forceinline__ double web_rate(double ci, int li, double fac, double bco, const double *sa)
{
    constexpr int G = GroupOf<NS>::G;
    double r = 0.0;
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        double cj = __shfl_sync(group_mask<G>(), ci, j, G);
        r = r + cj * sa[j * NS + (li < NS ? li : 0)];
    }
    return ci * (bco * fac + r);
}

// f:241-267 + res:213-220 at one point; returns delta for lane li
template <int NS, class State>
__device__ __forceinline__ double web_point_res(const WebDev &w, const State &S, i64 p, int li, double cdot_i,
                                                double ci, double fac, double bco, double cxi, double cyi,
                                                const double *sa)
{
    const int lic = li < NS ? li : 0;
    const i64 mxns = (i64)w.mx * NS;
    const int jyl = (int)(p / w.mx);          // local row, 0-based
    const int jx = (int)(p - (i64)jyl * w.mx);
    const int jyg = w.jy0 + jyl;              // global row, 0-based
    const i64 ici = p * NS + lic;
    // mirrored Neumann offsets (f:243-254); slab edges read the halo rows
    i64 idyu = mxns, idyl = mxns, idxu = NS, idxl = NS;
    if (jyg == w.my_g - 1) idyu = -mxns;
    if (jyg == 0) idyl = -mxns;
    if (jx == w.mx - 1) idxu = -NS;
    if (jx == 0) idxl = -NS;
    const double cyl = S(ici - idyl), cyu = S(ici + idyu), cxl = S(ici - idxl), cxu = S(ici + idxu);
    const double R = web_rate<NS>(ci, li, fac, bco, sa);
    const double dcyli = ci - cyl, dcyui = cyu - ci, dcxli = ci - cxl, dcxui = cxu - ci;
    const double f = cyi * (dcyui - dcyli) + cxi * (dcxui - dcxli) + R;
    return (lic >= w.np) ? -f : (cdot_i - f);
}

__device__ __forceinline__ double web_fac(const WebDev &w, i64 p)
{
    const int jyl = (int)(p / w.mx);
    const int jx = (int)(p - (i64)jyl * w.mx);
    const int jyg = w.jy0 + jyl;
    const double y = jyg * w.dy, x = jx * w.dx;
    return 1.0 + w.alpha * x * y + w.beta * __ldg(w.sx + jx) * __ldg(w.sy + jyg);
}

#define GROUP_LOOP_PROLOGUE(NS)                                                            \
    constexpr int G = GroupOf<NS>::G;                                                      \
    const int li = threadIdx.x % G;                                                        \
    const int lic = li < NS ? li : 0;                                                      \
    const i64 gpb = blockDim.x / G;                                                        \
    const i64 ngroups = (i64)gridDim.x * gpb;                                              \
    const i64 npts = (i64)w.mx * w.my;                                                     \
    const i64 nloop = (npts + ngroups - 1) / ngroups;                                      \
    extern __shared__ double sa[];                                                         \
    for (int t = threadIdx.x; t < NS * NS; t += blockDim.x) sa[t] = w.acoef[t];            \
    __syncthreads();                                                                       \
    const double bco = w.bcoef[lic], cxi = w.cox[lic], cyi = w.coy[lic];

// ---------------------------------------------------------------- K4: residual
template <int NS>
__global__ void __launch_bounds__(256) web_res_kernel(WebDev w, const double *__restrict__ c,
                                                      const double *__restrict__ cdot, double *__restrict__ delta)
{
    GROUP_LOOP_PROLOGUE(NS)
    StatePlain S{c};
    for (i64 it = 0; it < nloop; ++it) {
        i64 p = (i64)blockIdx.x * gpb + threadIdx.x / G + it * ngroups;
        const bool live = p < npts;
        if (!live) p = npts - 1;
        const i64 idx = p * NS + lic;
        const double ci = S(idx);
        const double d = web_point_res<NS>(w, S, p, li, __ldg(cdot + idx), ci, web_fac(w, p), bco, cxi, cyi, sa);
        if (live && li < NS) delta[idx] = d;
    }
}

// ---------------------------------------------------------------- K5: FD block Jacobian + LU
template <int NS>
__global__ void __launch_bounds__(128) web_jac_kernel(WebDev w, int nsd, const double *__restrict__ c,
                                                      const double *__restrict__ rewt, double cj,
                                                      double *__restrict__ bd, int *__restrict__ ipbd,
                                                      int *first_flag)
{
    GROUP_LOOP_PROLOGUE(NS)
    (void)cxi; (void)cyi;
    for (i64 it = 0; it < nloop; ++it) {
        i64 p = (i64)blockIdx.x * gpb + threadIdx.x / G + it * ngroups;
        const bool live = p < npts;
        if (!live) p = npts - 1;
        const i64 idx = p * NS + lic;
        const double ci = __ldg(c + idx);
        const double rw = __ldg(rewt + idx);
        const double fac = web_fac(w, p);
        const double r0 = web_rate<NS>(ci, li, fac, bco, sa);
        double a[NS];
#pragma unroll
        for (int js = 0; js < NS; ++js) {
            // jac_rbdpre:219-229: del = max(srur*|u_j|, 0.01/rewt_j); perturb; column = (r1-r0)*(-1/del)
            const double uj = __shfl_sync(group_mask<G>(), ci, js, G);
            const double rj = __shfl_sync(group_mask<G>(), rw, js, G);
            const double del = fmax(w.srur * fabs(uj), 1e-2 / rj);
            const double cpert = (li == js) ? (uj + del) : ci;
            const double fd = -1.0 / del;
            const double r1 = web_rate<NS>(cpert, li, fac, bco, sa);
            a[js] = (r1 - r0) * fd;
        }
        // add cj*I_d (jac_rbdpre:240-244)
#pragma unroll
        for (int j = 0; j < NS; ++j)
            if (j == li && li < nsd) a[j] = a[j] + cj;
        int ipv;
        const int inf = group_dgefa<NS>(a, li, ipv);
        if (live && li < NS) {
            // interleaved layout of psol_tpp.cu: element (i,j) -> ((p/32)*ns*ns + j*ns + i)*32 + p%32
            double *blk = bd + ((p >> 5) * (NS * NS)) * 32 + (p & 31);
#pragma unroll
            for (int j = 0; j < NS; ++j) blk[(j * NS + li) * 32] = a[j];
            ipbd[((p >> 5) * NS + li) * 32 + (p & 31)] = ipv;
            if (li == 0 && inf != 0) atomicMin(first_flag, (int)(p < 2147483646 ? p + 1 : 2147483646));
        }
    }
}

// ---------------------------------------------------------------- K3a: datv residual
// vtemp = G(t, y + v/wght, ydot + cj*v/wght) (datv:3493-3503) without materialising z / ydottemp;
// the block solve, "- savr" and "* wght" (datv:3509-3518) follow in tpp_psol (psol_tpp.cu).
template <int NS>
__global__ void __launch_bounds__(256) web_resdatv_kernel(WebDev w, const double *__restrict__ v,
                                                          const double *__restrict__ wt, double wscale,
                                                          const double *__restrict__ y, const double *__restrict__ ydot,
                                                          double cj, double *__restrict__ vtemp)
{
    GROUP_LOOP_PROLOGUE(NS)
    StateDatv S{y, v, wt, wscale};
    for (i64 it = 0; it < nloop; ++it) {
        i64 p = (i64)blockIdx.x * gpb + threadIdx.x / G + it * ngroups;
        const bool live = p < npts;
        if (!live) p = npts - 1;
        const i64 idx = p * NS + lic;
        const double wgt = wscale * __ldg(wt + idx);
        const double vtmp = __ldg(v + idx) / wgt;          // vtemp = v/wght
        const double zi = __ldg(y + idx) + vtmp;           // z = y + vtemp
        const double ydt = __ldg(ydot + idx) + vtmp * cj;  // ydottemp = ydot + vtemp*cj
        const double d = web_point_res<NS>(w, S, p, li, ydt, zi, web_fac(w, p), bco, cxi, cyi, sa);
        if (live && li < NS) vtemp[idx] = d;
    }
}

// ---------------------------------------------------------------- launchers
static int grp_grid(i64 npts, int G, int block, int per_sm)
{
    i64 gpb = block / G;
    i64 g = (npts + gpb - 1) / gpb;
    i64 cap = 148 * (i64)per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

#define WEB_DISPATCH(ns, CALL)                                                              \
    switch (ns) {                                                                           \
    case 2: { CALL(2); } break;                                                             \
    case 4: { CALL(4); } break;                                                             \
    case 6: { CALL(6); } break;                                                             \
    case 8: { CALL(8); } break;                                                             \
    case 10: { CALL(10); } break;                                                           \
    case 12: { CALL(12); } break;                                                           \
    case 16: { CALL(16); } break;                                                           \
    case 20: { CALL(20); } break;                                                           \
    case 24: { CALL(24); } break;                                                           \
    case 32: { CALL(32); } break;                                                           \
    default:                                                                                \
        fprintf(stderr, "daskr_b200: food-web ns=%d not instantiated\n", ns);               \
        abort();                                                                            \
    }

void web_res_launch(const double *c, const double *cdot, double *delta)
{
    DkbCtx *x = g_ctx;
    WebDev w = make_webdev();
    const i64 npts = x->npts;
    const size_t sh = sizeof(double) * w.ns * w.ns;
#define CALL(N) web_res_kernel<N><<<grp_grid(npts, GroupOf<N>::G, 256, 8), 256, sh, x->stream>>>(w, c, cdot, delta)
    WEB_DISPATCH(w.ns, CALL)
#undef CALL
    COUNT_LAUNCH();
}

void web_jac_launch(const double *c, const double *rewt, double cj, double *bd, int *ipbd, int *d_first_flag)
{
    DkbCtx *x = g_ctx;
    WebDev w = make_webdev();
    const i64 npts = x->npts;
    const size_t sh = sizeof(double) * w.ns * w.ns;
    const int nsd = x->nsd;
#define CALL(N) web_jac_kernel<N><<<grp_grid(npts, GroupOf<N>::G, 128, 8), 128, sh, x->stream>>>(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag)
    WEB_DISPATCH(w.ns, CALL)
#undef CALL
    COUNT_LAUNCH();
}

void web_datv_fused(const double *v, const double *wt, double wscale, const double *y, const double *ydot,
                    const double *savr, double cj, const double *bd, const int *ipbd, double *vnext)
{
    DkbCtx *x = g_ctx;
    WebDev w = make_webdev();
    const i64 npts = x->npts;
    const size_t sh = sizeof(double) * w.ns * w.ns;
    double *vtemp = x->z;
#define CALL(N) web_resdatv_kernel<N><<<grp_grid(npts, GroupOf<N>::G, 256, 8), 256, sh, x->stream>>>(w, v, wt, wscale, y, ydot, cj, vtemp)
    WEB_DISPATCH(w.ns, CALL)
#undef CALL
    COUNT_LAUNCH();
    tpp_psol(bd, ipbd, w.ns, npts, vtemp, savr, wt, wscale, vnext, -1);
}
