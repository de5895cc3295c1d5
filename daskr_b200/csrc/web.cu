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

#include "web_dev.cuh"

WebDev make_webdev()
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
    w.dy = c->web.dyn;                 // normalised row spacing (== dy for the unit square)
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
    for (int jy = 1; jy <= c->my_g; ++jy) tab.push_back(sin(4 * pi * ((jy - 1) * p.dyn)));
    for (int i = 0; i < p.ns * p.ns; ++i) tab.push_back(p.acoef[i]);
    for (int i = 0; i < p.ns; ++i) tab.push_back(p.bcoef[i]);
    for (int i = 0; i < p.ns; ++i) tab.push_back(p.cox[i]);
    for (int i = 0; i < p.ns; ++i) tab.push_back(p.coy[i]);
    if (c->d_sx) cudaFree(c->d_sx);
    cudaMalloc(&c->d_sx, tab.size() * sizeof(double));
    cudaMemcpy(c->d_sx, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice);
    datv_fused_upload_tables(p);
    jac2_upload_tables(p);
    dkb_cuda_check("web_upload_tables");
}

__device__ __forceinline__ double web_fac_xy(const WebDev &w, int jx, int jyg)
{
    const double y = jyg * w.dy, x = jx * w.dx;
    return 1.0 + w.alpha * x * y + w.beta * __ldg(w.sx + jx) * __ldg(w.sy + jyg);
}

// ---------------------------------------------------------------- K4 / K3a: tiled residual
// Tile = RT<NS>::TX mesh points wide, RT_TY rows high; one thread per (point-in-row, species),
// i.e. NS*TX threads, each walking the RT_TY rows of the tile with its species fixed -- so its
// row of acoef, bcoef, cox, coy live in registers and the interaction sum costs one (broadcast)
// shared-memory read per term.
#define RT_TY 8
template <int NS> struct RT {
    static constexpr int TX = NS >= 16 ? 16 : NS >= 8 ? 32 : NS >= 4 ? 64 : 128;
    static constexpr int THREADS = NS * TX;
};

template <int NS, bool DATV>
__global__ void __launch_bounds__(RT<NS>::THREADS)
web_restile_kernel(WebDev w, const double *__restrict__ v, double vscale, const double *__restrict__ wt, double wscale,
                   const double *__restrict__ y, const double *__restrict__ ydot, double cj,
                   const double *__restrict__ sub, double *__restrict__ out)
{
    // sub != nullptr: out = G - sub (datv's "vtemp - savr", src/daskr_new.f90:3509, folded into the store)
    constexpr int TX = RT<NS>::TX, THREADS = RT<NS>::THREADS;
    constexpr int ZX = TX + 2;
    constexpr int NHALO = (2 * TX + 2 * RT_TY) * NS;
    extern __shared__ double Z[];               // state values, [RT_TY+2][ZX][NS]
    const int tid = threadIdx.x;
    const int x = tid / NS, i = tid - x * NS;   // this thread's point within a tile row, its species
    double arow[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) arow[j] = __ldg(w.acoef + j * NS + i);
    const double bco = __ldg(w.bcoef + i), cxi = __ldg(w.cox + i), cyi = __ldg(w.coy + i);

    const int tiles_x = (w.mx + TX - 1) / TX;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int x0 = tx * TX, y0 = ty * RT_TY;
    const int gx = x0 + x;
    const i64 mxns = (i64)w.mx * NS;

    // ---- phase 1a: this thread's own unknowns, one per tile row (kept in registers across the barrier).
    // Loads are unconditional (cells outside the slab read a safe in-tile address and are discarded)
    // so that all of them are in flight before the first divide.
    double ydt[RT_TY];
    {
        double rw[RT_TY], rv[RT_TY], ryy[RT_TY];
        const i64 gsafe = (i64)y0 * mxns + (i64)x0 * NS;
#pragma unroll
        for (int ry = 0; ry < RT_TY; ++ry) {
            const int gy = y0 + ry;
            // rows of the tile beyond the slab: the first one is the upper neighbour's ghost row and
            // is needed by the last local row when the slab height is not a multiple of the tile height
            const bool ok = (gy < w.my || (gy == w.my && w.has_hi)) && gx < w.mx;
            const i64 g = ok ? ((i64)gy * mxns + (i64)gx * NS + i) : gsafe;
            ryy[ry] = __ldg(y + g);
            ydt[ry] = __ldg(ydot + g);
            if (DATV) {
                rw[ry] = __ldg(wt + g);
                rv[ry] = __ldg(v + g);
            }
        }
#pragma unroll
        for (int ry = 0; ry < RT_TY; ++ry) {
            const int gy = y0 + ry;
            const bool ok = (gy < w.my || (gy == w.my && w.has_hi)) && gx < w.mx;
            double z = ryy[ry];
            if (DATV) {
                const double wgt = wscale * rw[ry];
                const double vtmp = (vscale * rv[ry]) / wgt;   // vtemp = v/wght            (datv:3493)
                z = ryy[ry] + vtmp;                            // z = y + vtemp             (:3498)
                ydt[ry] = ydt[ry] + vtmp * cj;                 // ydottemp = ydot + vtemp*cj (:3497)
            }
            if (ok) Z[((ry + 1) * ZX + (x + 1)) * NS + i] = z;
        }
    }
    // ---- phase 1b: the stencil halo of the tile (no corners)
    for (int h = tid; h < NHALO; h += THREADS) {
        const int hp = h / NS, hi = h - hp * NS;
        int ry, hx;
        if (hp < TX) { ry = -1; hx = hp; }
        else if (hp < 2 * TX) { ry = RT_TY; hx = hp - TX; }
        else if (hp < 2 * TX + RT_TY) { ry = hp - 2 * TX; hx = -1; }
        else { ry = hp - 2 * TX - RT_TY; hx = TX; }
        const int gy = y0 + ry, hgx = x0 + hx;
        const bool top_bottom = (ry < 0 || ry >= RT_TY);
        const bool rowok = top_bottom ? ((gy >= 0 && gy < w.my) || (gy == -1 && w.has_lo) || (gy == w.my && w.has_hi))
                                      : (gy < w.my);
        if (rowok && hgx >= 0 && hgx < w.mx) {
            const i64 g = (i64)gy * mxns + (i64)hgx * NS + hi;   // may address the slab halo rows
            double z = __ldg(y + g);
            if (DATV) z = z + (vscale * __ldg(v + g)) / (wscale * __ldg(wt + g));
            Z[((ry + 1) * ZX + (hx + 1)) * NS + hi] = z;
        }
    }
    __syncthreads();

    // ---- phase 2: f:241-267, rates:286-297, res:213-220
    if (gx >= w.mx) return;
    const int dxl = (gx == 0) ? 1 : -1;           // mirrored Neumann neighbours (f:243-254)
    const int dxu = (gx == w.mx - 1) ? -1 : 1;
    const double sxv = __ldg(w.sx + gx);
    const double xx = gx * w.dx;
#pragma unroll
    for (int ry = 0; ry < RT_TY; ++ry) {
        const int gy = y0 + ry;
        if (gy >= w.my) break;
        const int jyg = w.jy0 + gy;
        const int pt = (ry + 1) * ZX + (x + 1);
        const int pyl = (jyg == 0) ? pt + ZX : pt - ZX;
        const int pyu = (jyg == w.my_g - 1) ? pt - ZX : pt + ZX;
        const double *zp = Z + pt * NS;
        const double ci = zp[i];
        double r = 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j) r = r + zp[j] * arow[j];
        const double yy = jyg * w.dy;
        const double fac = 1.0 + w.alpha * xx * yy + w.beta * sxv * __ldg(w.sy + jyg);
        const double R = ci * (bco * fac + r);
        const double dcyli = ci - Z[pyl * NS + i], dcyui = Z[pyu * NS + i] - ci;
        const double dcxli = ci - Z[(pt + dxl) * NS + i], dcxui = Z[(pt + dxu) * NS + i] - ci;
        const double f = cyi * (dcyui - dcyli) + cxi * (dcxui - dcxli) + R;
        const i64 o = (i64)gy * mxns + (i64)gx * NS + i;
        const double g = (i >= w.np) ? -f : (ydt[ry] - f);
        out[o] = sub ? g - sub[o] : g;
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
    i64 cap = g_ctx->nsm * (i64)per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// ns = 2*np: every even block size up to 32
#define WEB_DISPATCH(ns, CALL) \
    switch (ns) { \
    case 2: { CALL(2); } break; \
    case 4: { CALL(4); } break; \
    case 6: { CALL(6); } break; \
    case 8: { CALL(8); } break; \
    case 10: { CALL(10); } break; \
    case 12: { CALL(12); } break; \
    case 14: { CALL(14); } break; \
    case 16: { CALL(16); } break; \
    case 18: { CALL(18); } break; \
    case 20: { CALL(20); } break; \
    case 22: { CALL(22); } break; \
    case 24: { CALL(24); } break; \
    case 26: { CALL(26); } break; \
    case 28: { CALL(28); } break; \
    case 30: { CALL(30); } break; \
    case 32: { CALL(32); } break; \
    default: \
        fprintf(stderr, "daskr_b200: food-web ns=%d out of range (even, 2..32)\n", ns); \
        abort(); \
    }

template <int NS, bool DATV>
static void restile_launch(const WebDev &w, const double *v, double vscale, const double *wt, double wscale, const double *y,
                           const double *ydot, double cj, const double *sub, double *out)
{
    static bool attr_set = false;
    constexpr int TX = RT<NS>::TX;
    const size_t sh = sizeof(double) * ((TX + 2) * (RT_TY + 2) * NS);
    if (!attr_set && sh > 48 * 1024) {
        cudaFuncSetAttribute(web_restile_kernel<NS, DATV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
        attr_set = true;
    }
    const int tiles = ((w.mx + TX - 1) / TX) * ((w.my + RT_TY - 1) / RT_TY);
    web_restile_kernel<NS, DATV><<<tiles, RT<NS>::THREADS, sh, g_ctx->stream>>>(w, v, vscale, wt, wscale, y, ydot, cj, sub, out);
    COUNT_LAUNCH();
}

void web_res_launch(const double *c, const double *cdot, double *delta)
{
    WebDev w = make_webdev();
    if (g_ctx->opt_fused >= 2 && web_resid_single(w, 2, nullptr, 1.0, nullptr, 1.0, c, cdot, nullptr, 0.0, delta)) return;
#define CALL(N) restile_launch<N, false>(w, nullptr, 1.0, nullptr, 1.0, c, cdot, 0.0, nullptr, delta)
    WEB_DISPATCH(w.ns, CALL)
#undef CALL
}

bool web_jac_launch(const double *c, const double *rewt, double cj, double *bd, int *ipbd, int *d_first_flag, bool exact)
{
    DkbCtx *x = g_ctx;
    WebDev w = make_webdev();
    const i64 npts = x->npts;
    const int nsd = x->nsd;
    // debug_variant 9: the one-mapping kernel below; 10..15: jac2.cu (its own variants 0..5 shifted by 10)
    if (!exact && (x->debug_variant < 9 || x->debug_variant > 15) && web_jac3(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag)) return true;   // two rows per lane (jac3.cu)
    if (x->debug_variant != 9 && web_jac2(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag)) return false;   // two-phase kernel (jac2.cu)
#define CALL(N) web_jac_kernel<N><<<grp_grid(npts, GroupOf<N>::G, 128, 8), 128, 0, x->stream>>>(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag)
    WEB_DISPATCH(w.ns, CALL)
#undef CALL
    COUNT_LAUNCH();
    return false;
}

// ---------------------------------------------------------------- block grouping (jbg = 1)
// jac_rbgpre, src/daskr_rbgpre.f90:195-254: one FD block per GROUP, evaluated at the group's representative point.
// Same arithmetic as web_jac_kernel; the block is written unfactored in the reference layout (group g at g*ns^2,
// column-major) -- zeros if the representative point lives on another rank -- so that the ranks can sum their
// parts before cj*I_d is added and the blocks are factored.
template <int NS>
__global__ void __launch_bounds__(128) web_jac_groups_kernel(WebDev w, const i64 *__restrict__ rep, int ngrp,
                                                             const double *__restrict__ c,
                                                             const double *__restrict__ rewt, double *__restrict__ bd)
{
    constexpr int G = GroupOf<NS>::G;
    const unsigned mask = group_mask<G>();
    const int li = threadIdx.x % G;
    const int lic = li < NS ? li : 0;
    const int gpb = blockDim.x / G;
    double arow[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) arow[j] = w.acoef[j * NS + lic];
    const double bco = w.bcoef[lic];
    const int nloop = (ngrp + gridDim.x * gpb - 1) / (gridDim.x * gpb);
    for (int it = 0; it < nloop; ++it) {
        int g = blockIdx.x * gpb + threadIdx.x / G + it * gridDim.x * gpb;
        const bool live = g < ngrp;
        if (!live) g = ngrp - 1;
        const i64 pr = rep[g];
        const bool own = pr >= 0;
        const i64 p = own ? pr : 0;
        const i64 idx = p * NS + lic;
        const double ci = __ldg(c + idx);
        const double rw = __ldg(rewt + idx);
        const int jyl = (int)(p / w.mx);
        const int jx = (int)(p - (i64)jyl * w.mx);
        const double fac = web_fac_xy(w, jx, w.jy0 + jyl);
        double cb[NS], del[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            cb[j] = __shfl_sync(mask, ci, j, G);
            const double rj = __shfl_sync(mask, rw, j, G);
            del[j] = fmax(w.srur * fabs(cb[j]), 1e-2 / rj);   // jac_rbgpre:233
        }
        double r = 0.0;
#pragma unroll
        for (int j = 0; j < NS; ++j) r = r + cb[j] * arow[j];
        const double r0 = ci * (bco * fac + r);
#pragma unroll
        for (int js = 0; js < NS; ++js) {
            const double up = cb[js] + del[js];
            const double fd = -1.0 / del[js];
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < NS; ++j) s = s + ((j == js) ? up : cb[j]) * arow[j];
            const double cown = (li == js) ? up : ci;
            const double r1 = cown * (bco * fac + s);
            if (live && li < NS) bd[(i64)g * (NS * NS) + js * NS + li] = own ? (r1 - r0) * fd : 0.0;
        }
    }
}

void web_jac_groups_launch(const double *c, const double *rewt, double *bd)
{
    DkbCtx *x = g_ctx;
    WebDev w = make_webdev();
    const int ngrp = x->ngx * x->ngy;
#define CALL(N) web_jac_groups_kernel<N><<<(ngrp * GroupOf<N>::G + 127) / 128, 128, 0, x->stream>>>(w, x->d_rep, ngrp, c, rewt, bd)
    WEB_DISPATCH(w.ns, CALL)
#undef CALL
    COUNT_LAUNCH();
}

// datv (src/daskr_new.f90:3429-3520) for the built-in problem: vtemp = G(z, ydottemp) into the
// ydottemp slot (free here: ydottemp is never materialised), then the block solve with the
// "- savr" and "* wght" statements folded into its load / store phases.
// jpre = 3 (example_web.f90:346-391, the shipped setting): the Gauss-Seidel factor runs between the two, on
// vtemp - savr (the subtraction folded into the residual kernel's store).
int web_datv_fused(const double *v, double vscale, const double *wt, double wscale, const double *y, const double *ydot,
                   const double *savr, double cj, const double *bd, const int *ipbd, double *vnext, int jpre)
{
    DkbCtx *x = g_ctx;
    WebDev w = make_webdev();
    // one kernel when the mesh allows it (warp = 32 consecutive points of a row = one LU tile)
    if (jpre == 1 && x->opt_fused >= 2 && web_datv_single(w, v, vscale, wt, wscale, y, ydot, savr, cj, bd, ipbd, vnext)) return 0;
    double *vtemp = x->z;
    const double *sub = (jpre == 1) ? nullptr : savr;
    {
        ProfScope ps(PROF_RESDATV);
        if (!(jpre != 1 && x->opt_fused >= 2 && web_resid_single(w, 1, v, vscale, wt, wscale, y, ydot, savr, cj, vtemp))) {
#define CALL(N) restile_launch<N, true>(w, v, vscale, wt, wscale, y, ydot, cj, sub, vtemp)
            WEB_DISPATCH(w.ns, CALL)
#undef CALL
        }
    }
    double *gsout = vtemp;   // several ranks: the factor may leave its result in the scratch vector (gs.cu)
    if (jpre != 1 && web_gauss_seidel_to(1.0 / cj, vtemp, x->wk, &gsout) != 0) return -1;
    return tpp_psol(bd, ipbd, w.ns, x->npts, gsout, (jpre == 1) ? savr : nullptr, wt, wscale, vnext, -1) != 0 ? -1 : 0;
}
