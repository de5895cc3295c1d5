// web.cu -- device kernels of the food-web problem (example/example_web.f90)
//   K4/K3a  residual        res:190-224 + f:226-269 + rates:271-299, evaluated either at (c, cdot)
//                           or, for datv (src/daskr_new.f90:3493-3503), at the perturbed state
//                           z = y + v/wght, ydottemp = ydot + cj*v/wght formed on the fly
//   K5      block Jacobian  jac:301-344 -> jac_rbdpre (src/daskr_rbdpre.f90:158-251) + dgefa
// The block solve that completes datv (psol_rbdpre, "- savr", "* wght") is psol_tpp.cu.
//
// Residual: one CTA per 16 x 8 tile of mesh points, one thread per unknown (species fastest, so
// every global access is a coalesced 2560-byte row segment).  The state values of the tile and of
// its 5-point-stencil halo are staged once in shared memory -- for datv that includes the divide
// v/wght, done 1.375x per unknown instead of 5x -- and both the stencil and the ns x ns interaction
// sum read them from there.  Mesh rows are slab-decomposed: rows outside the slab come from the
// halo rows stored just before / after the local array, global edges use the reference's mirrored
// Neumann differences.
//
// Jacobian: one sub-warp group per mesh point, lane i = row i of the block.  The point's
// concentrations and this lane's row of acoef are hoisted into registers, so the ns finite-
// difference columns need no shuffles; the LU factorisation keeps rows where they are and tracks
// LINPACK's row interchanges as a logical permutation (lu_device.cuh), which leaves one shuffle
// broadcast per eliminated column.  Arithmetic is the reference's, operation for operation.
#include "dkb_internal.h"
#include <math.h>
#include <string.h>

#include "lu_device.cuh"   // group_dgefa_track

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

__device__ __forceinline__ double web_fac_xy(const WebDev &w, int jx, int jyg)
{
    const double y = jyg * w.dy, x = jx * w.dx;
    return 1.0 + w.alpha * x * y + w.beta * __ldg(w.sx + jx) * __ldg(w.sy + jyg);
}

// ---------------------------------------------------------------- K4 / K3a: tiled residual
#define RT_TX 16
#define RT_TY 8
#define RT_THREADS 256

template <int NS, bool DATV>
__global__ void __launch_bounds__(RT_THREADS)
web_restile_kernel(WebDev w, const double *__restrict__ v, const double *__restrict__ wt, double wscale,
                   const double *__restrict__ y, const double *__restrict__ ydot, double cj,
                   double *__restrict__ out)
{
    constexpr int ZX = RT_TX + 2, ZY = RT_TY + 2;
    constexpr int ROWU = RT_TX * NS;            // unknowns per tile row
    constexpr int NU = RT_TY * ROWU;            // unknowns per tile
    constexpr int PER = (NU + RT_THREADS - 1) / RT_THREADS;
    constexpr int NHALO = (2 * RT_TX + 2 * RT_TY) * NS;
    extern __shared__ double sm[];
    double *sa = sm;                            // acoef, column-major
    double *sb = sa + NS * NS;                  // bcoef | cox | coy
    double *Z = sb + 3 * NS;                    // state values, [ZY][ZX][NS]
    const int tid = threadIdx.x;
    for (int t = tid; t < NS * NS; t += RT_THREADS) sa[t] = w.acoef[t];
    for (int t = tid; t < 3 * NS; t += RT_THREADS) sb[t] = w.bcoef[t];   // bcoef, cox, coy are contiguous

    const int tiles_x = (w.mx + RT_TX - 1) / RT_TX;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int x0 = tx * RT_TX, y0 = ty * RT_TY;
    const i64 mxns = (i64)w.mx * NS;

    // state value at global unknown g (may index the halo rows before/after the local array)
    auto state = [&](i64 g) -> double {
        if (DATV) return __ldg(y + g) + __ldg(v + g) / (wscale * __ldg(wt + g));   // z = y + v/wght
        return __ldg(y + g);
    };

    // ---- phase 1a: this thread's own unknowns (kept in registers across the barrier)
    double ydt[PER];
#pragma unroll
    for (int m = 0; m < PER; ++m) {
        const int q = tid + RT_THREADS * m;
        ydt[m] = 0.0;
        if (q < NU) {
            const int ry = q / ROWU, rem = q - ry * ROWU;
            const int x = rem / NS, i = rem - x * NS;
            const int gy = y0 + ry, gx = x0 + x;
            if (gy < w.my && gx < w.mx) {
                const i64 g = (i64)gy * mxns + (i64)gx * NS + i;
                double z;
                if (DATV) {
                    const double wgt = wscale * __ldg(wt + g);
                    const double vtmp = __ldg(v + g) / wgt;        // vtemp = v/wght        (datv:3493)
                    z = __ldg(y + g) + vtmp;                       // z = y + vtemp         (:3498)
                    ydt[m] = __ldg(ydot + g) + vtmp * cj;          // ydottemp = ydot+vtemp*cj (:3497)
                } else {
                    z = __ldg(y + g);
                    ydt[m] = __ldg(ydot + g);
                }
                Z[((ry + 1) * ZX + (x + 1)) * NS + i] = z;
            }
        }
    }
    // ---- phase 1b: the stencil halo of the tile (no corners)
    for (int h = tid; h < NHALO; h += RT_THREADS) {
        const int hp = h / NS, i = h - hp * NS;
        int ry, x;
        if (hp < RT_TX) { ry = -1; x = hp; }
        else if (hp < 2 * RT_TX) { ry = RT_TY; x = hp - RT_TX; }
        else if (hp < 2 * RT_TX + RT_TY) { ry = hp - 2 * RT_TX; x = -1; }
        else { ry = hp - 2 * RT_TX - RT_TY; x = RT_TX; }
        const int gy = y0 + ry, gx = x0 + x;
        const bool rowok = (gy >= 0 && gy < w.my) || (gy == -1 && w.has_lo) || (gy == w.my && w.has_hi);
        if (rowok && gx >= 0 && gx < w.mx && (ry < 0 || ry >= RT_TY || gy < w.my))
            Z[((ry + 1) * ZX + (x + 1)) * NS + i] = state((i64)gy * mxns + (i64)gx * NS + i);
    }
    __syncthreads();

    // ---- phase 2: f:241-267, rates:286-297, res:213-220
#pragma unroll
    for (int m = 0; m < PER; ++m) {
        const int q = tid + RT_THREADS * m;
        if (q >= NU) continue;
        const int ry = q / ROWU, rem = q - ry * ROWU;
        const int x = rem / NS, i = rem - x * NS;
        const int gy = y0 + ry, gx = x0 + x;
        if (gy >= w.my || gx >= w.mx) continue;
        const int jyg = w.jy0 + gy;
        const int pt = (ry + 1) * ZX + (x + 1);
        // mirrored Neumann neighbours (f:243-254)
        const int pyl = (jyg == 0) ? pt + ZX : pt - ZX;
        const int pyu = (jyg == w.my_g - 1) ? pt - ZX : pt + ZX;
        const int pxl = (gx == 0) ? pt + 1 : pt - 1;
        const int pxu = (gx == w.mx - 1) ? pt - 1 : pt + 1;
        const double *zp = Z + pt * NS;
        const double ci = zp[i];
        double r = 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j) r = r + zp[j] * sa[j * NS + i];
        const double fac = web_fac_xy(w, gx, jyg);
        const double R = ci * (sb[i] * fac + r);
        const double dcyli = ci - Z[pyl * NS + i], dcyui = Z[pyu * NS + i] - ci;
        const double dcxli = ci - Z[pxl * NS + i], dcxui = Z[pxu * NS + i] - ci;
        const double f = sb[2 * NS + i] * (dcyui - dcyli) + sb[NS + i] * (dcxui - dcxli) + R;
        out[(i64)gy * mxns + (i64)gx * NS + i] = (i >= w.np) ? -f : (ydt[m] - f);
    }
}

// ---------------------------------------------------------------- K5: FD block Jacobian + LU
template <int NS>
__global__ void __launch_bounds__(128) web_jac_kernel(WebDev w, int nsd, const double *__restrict__ c,
                                                      const double *__restrict__ rewt, double cj,
                                                      double *__restrict__ bd, int *__restrict__ ipbd,
                                                      int *first_flag)
{
    constexpr int G = GroupOf<NS>::G;
    const unsigned mask = group_mask<G>();
    const int li = threadIdx.x % G;
    const int lic = li < NS ? li : 0;
    const i64 gpb = blockDim.x / G;
    const i64 ngroups = (i64)gridDim.x * gpb;
    const i64 npts = (i64)w.mx * w.my;
    const i64 nloop = (npts + ngroups - 1) / ngroups;
    // this lane's row of acoef and its bcoef: rates:290-297 needs sum_j c_j*acoef(i,j)
    double arow[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) arow[j] = w.acoef[j * NS + lic];
    const double bco = w.bcoef[lic];

    for (i64 it = 0; it < nloop; ++it) {
        i64 p = (i64)blockIdx.x * gpb + threadIdx.x / G + it * ngroups;
        const bool live = p < npts;
        if (!live) p = npts - 1;
        const i64 idx = p * NS + lic;
        const double ci = __ldg(c + idx);
        const double rw = __ldg(rewt + idx);
        const int jyl = (int)(p / w.mx);
        const int jx = (int)(p - (i64)jyl * w.mx);
        const double fac = web_fac_xy(w, jx, w.jy0 + jyl);
        // the point's concentrations and weights, broadcast once
        double cb[NS], del[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            cb[j] = __shfl_sync(mask, ci, j, G);
            const double rj = __shfl_sync(mask, rw, j, G);
            del[j] = fmax(w.srur * fabs(cb[j]), 1e-2 / rj);   // jac_rbdpre:222
        }
        double r = 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j) r = r + cb[j] * arow[j];
        const double r0 = ci * (bco * fac + r);               // what res left in rpar (f:250)
        double a[NS];
#pragma unroll
        for (int js = 0; js < NS; ++js) {
            // jac_rbdpre:219-229: perturb u_js, column = (r1 - r0)*(-1/del), restore
            const double up = cb[js] + del[js];
            const double fd = -1.0 / del[js];
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < NS; ++j) s = s + ((j == js) ? up : cb[j]) * arow[j];
            const double cown = (li == js) ? up : ci;
            const double r1 = cown * (bco * fac + s);
            a[js] = (r1 - r0) * fd;
        }
        // add cj*I_d (jac_rbdpre:240-244)
#pragma unroll
        for (int j = 0; j < NS; ++j)
            if (j == li && li < nsd) a[j] = a[j] + cj;
        // LU with logical row interchanges; column k is final after step k and is stored at its
        // LINPACK position pos right away (interleaved layout of psol_tpp.cu)
        double *blk = bd + ((p >> 5) * (NS * NS)) * 32 + (p & 31);
        int ipv;
        const int inf = group_dgefa_track<NS>(a, li, ipv, [&](int col, int pos, double val) {
            if (live && li < NS) blk[(col * NS + pos) * 32] = val;
        });
        if (live && li < NS) {
            ipbd[((p >> 5) * NS + li) * 32 + (p & 31)] = ipv;
            if (li == 0 && inf != 0) atomicMin(first_flag, (int)(p < 2147483646 ? p + 1 : 2147483646));
        }
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

template <int NS, bool DATV>
static void restile_launch(const WebDev &w, const double *v, const double *wt, double wscale, const double *y,
                           const double *ydot, double cj, double *out)
{
    static bool attr_set = false;
    const size_t sh = sizeof(double) * (NS * NS + 3 * NS + (RT_TX + 2) * (RT_TY + 2) * NS);
    if (!attr_set && sh > 48 * 1024) {
        cudaFuncSetAttribute(web_restile_kernel<NS, DATV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
        attr_set = true;
    }
    const int tiles = ((w.mx + RT_TX - 1) / RT_TX) * ((w.my + RT_TY - 1) / RT_TY);
    web_restile_kernel<NS, DATV><<<tiles, RT_THREADS, sh, g_ctx->stream>>>(w, v, wt, wscale, y, ydot, cj, out);
    COUNT_LAUNCH();
}

void web_res_launch(const double *c, const double *cdot, double *delta)
{
    WebDev w = make_webdev();
#define CALL(N) restile_launch<N, false>(w, nullptr, nullptr, 1.0, c, cdot, 0.0, delta)
    WEB_DISPATCH(w.ns, CALL)
#undef CALL
}

void web_jac_launch(const double *c, const double *rewt, double cj, double *bd, int *ipbd, int *d_first_flag)
{
    DkbCtx *x = g_ctx;
    WebDev w = make_webdev();
    const i64 npts = x->npts;
    const int nsd = x->nsd;
#define CALL(N) web_jac_kernel<N><<<grp_grid(npts, GroupOf<N>::G, 128, 8), 128, 0, x->stream>>>(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag)
    WEB_DISPATCH(w.ns, CALL)
#undef CALL
    COUNT_LAUNCH();
}

// datv (src/daskr_new.f90:3429-3520) for the built-in problem: vtemp = G(z, ydottemp) into the
// ydottemp slot (free here: ydottemp is never materialised), then the block solve with the
// "- savr" and "* wght" statements folded into its load / store phases.
void web_datv_fused(const double *v, const double *wt, double wscale, const double *y, const double *ydot,
                    const double *savr, double cj, const double *bd, const int *ipbd, double *vnext)
{
    DkbCtx *x = g_ctx;
    WebDev w = make_webdev();
    double *vtemp = x->z;
#define CALL(N) restile_launch<N, true>(w, v, wt, wscale, y, ydot, cj, vtemp)
    WEB_DISPATCH(w.ns, CALL)
#undef CALL
    tpp_psol(bd, ipbd, w.ns, x->npts, vtemp, savr, wt, wscale, vnext, -1);
}
