// jac2.cu -- K5: jac_rbdpre (src/daskr_rbdpre.f90:158-251) for the food-web problem in two phases
// inside one kernel, so that each phase can use the mapping that suits it:
//
//   phase FD (thread per (mesh point, block row), all 32 lanes of every warp busy):
//       bd(k,js) = -(R_k(u + del_js e_js) - R_k(u))/del_js, + cj on the first nsd diagonals,
//       with the point's concentrations, del and -1/del staged in shared memory and this thread's
//       row of acoef in registers.  The unfactored ns x ns blocks of the CTA's points land in shared
//       memory (rows conflict-free on write, columns conflict-free on read).
//   phase LU (one sub-warp group per point, lane = row): LINPACK dgefa with logical row
//       interchanges (group_dgefa_track, lu_device.cuh); every finished column goes straight to
//       the tile-interleaved global layout of psol_tpp.cu at its LINPACK position.
//
// The earlier single-mapping kernel (lane = row for both phases) ran the FD part on 20 of 32 lanes
// with 230 registers per thread (8 warps/SM): 10 ms per call at 1024^2 x 20.  Arithmetic is
// unchanged (reference operation order, no FMA), so blocks and pivots stay bit-identical to the
// CPU path.
#include "web_dev.cuh"
#include "lu_device.cuh"

template <int NS> struct J2 {
    static constexpr int PTS = NS >= 16 ? 16 : NS >= 8 ? 32 : NS >= 4 ? 64 : 128;   // default points per CTA
};

__constant__ double j_acoef[DKB_NS_MAX * DKB_NS_MAX];   // column-major (k,j) at [j*ns+k]
__constant__ double j_bcoef[DKB_NS_MAX];

void jac2_upload_tables(const WebPar &p)
{
    cudaMemcpyToSymbol(j_acoef, p.acoef, sizeof(double) * p.ns * p.ns);
    cudaMemcpyToSymbol(j_bcoef, p.bcoef, sizeof(double) * p.ns);
}

template <int NS, int PTS, bool SPEC>
__global__ void __launch_bounds__(NS * PTS)
web_jac2_kernel(WebDev w, int nsd, const double *__restrict__ c, const double *__restrict__ rewt, double cj,
                double *__restrict__ bd, int *__restrict__ ipbd, int *first_flag)
{
    constexpr int THREADS = NS * PTS;
    constexpr int G = GroupOf<NS>::G;
    constexpr int NSQ = NS * NS;
    constexpr int CP = NS + 1;              // column pitch in shared memory (odd: columns written by
    constexpr int BP = NS * CP;             // neighbouring threads fall into different banks)
    extern __shared__ double sm[];
    double *blk = sm;                       // [PTS][NS columns][CP]
    double *sc = blk + PTS * BP;            // [PTS][NS]    concentrations
    double *sr0 = sc + PTS * NS;            // [PTS][NS]    R(u), what res left in rpar (f:250)
    const int tid = threadIdx.x;
    const i64 npts = (i64)w.mx * w.my;
    const i64 p0 = (i64)blockIdx.x * PTS;

    // ---------------- phase FD: thread (point pl, column js) -- one perturbation per thread
    {
        const int pl = tid / NS, js = tid - pl * NS;
        const i64 p = p0 + pl;
        const i64 pp = p < npts ? p : npts - 1;
        const i64 idx = pp * NS + js;
        const double cjs = __ldg(c + idx);
        const double rw = __ldg(rewt + idx);
        sc[pl * NS + js] = cjs;
        __syncthreads();
        double cb[NS];
#pragma unroll
        for (int j = 0; j < NS; ++j) cb[j] = sc[pl * NS + j];
        const int jyl = (int)(pp / w.mx);
        const int jx = (int)(pp - (i64)jyl * w.mx);
        const int jyg = w.jy0 + jyl;
        const double yy = jyg * w.dy, xx = jx * w.dx;
        const double fac = 1.0 + w.alpha * xx * yy + w.beta * __ldg(w.sx + jx) * __ldg(w.sy + jyg);
        // R_js(u) by this thread (rates:290-297 for row js), shared with the point's other threads
        {
            double r = 0.0;
#pragma unroll
            for (int j = 0; j < NS; ++j) r = r + cb[j] * __ldg(w.acoef + j * NS + js);
            sr0[pl * NS + js] = cjs * (__ldg(w.bcoef + js) * fac + r);
        }
        const double del = fmax(w.srur * fabs(cjs), 1e-2 / rw);   // jac_rbdpre:222
        const double up = cjs + del;                              // u(j) = u(j) + del  (:223)
        const double fd = -1.0 / del;                             // fac = -one/del     (:224)
#pragma unroll
        for (int j = 0; j < NS; ++j)
            if (j == js) cb[j] = up;                              // the perturbed point state
        __syncthreads();
        double *bp = blk + pl * BP + js * CP;
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            // rblock at the perturbed state, row k (acoef is a uniform constant-bank operand here)
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < NS; ++j) s = s + cb[j] * j_acoef[j * NS + k];
            const double r1 = cb[k] * (j_bcoef[k] * fac + s);
            double val = (r1 - sr0[pl * NS + k]) * fd;            // (:226-228)
            if (k == js && k < nsd) val = val + cj;               // + cj*I_d (:240-244)
            bp[k] = val;
        }
    }
    __syncthreads();

    // ---------------- phase LU
    {
        const int li = tid % G, gid = tid / G;
        constexpr int NGRP = THREADS / G;
        constexpr int ROUNDS = (PTS + NGRP - 1) / NGRP;
#pragma unroll 1
        for (int rd = 0; rd < ROUNDS; ++rd) {
            const int pl = gid + rd * NGRP;
            const i64 p = p0 + pl;
            const bool live = (pl < PTS) && (p < npts);
            const int plc = pl < PTS ? pl : PTS - 1;
            double a[NS];
#pragma unroll
            for (int j = 0; j < NS; ++j) a[j] = (li < NS) ? blk[plc * BP + j * CP + li] : 0.0;
            const i64 pq = live ? p : 0;
            double *gb = bd + ((pq >> 5) * NSQ) * 32 + (pq & 31);
            int ipv;
            // (broadcasting the pivot row through shared memory -- 128-bit stores by the pivot lane, 128-bit broadcast
            // loads -- instead of shuffles was measured at 8.2 ms against 5.4 ms per call at 1024^2 x 20)
            const int inf = group_dgefa_track<NS, SPEC>(a, li, ipv, [&](int col, int pos, double val) {
                if (live && li < NS) gb[(col * NS + pos) * 32] = val;
            });
            if (live && li < NS) {
                ipbd[((p >> 5) * NS + li) * 32 + (p & 31)] = ipv;
                if (li == 0 && inf != 0) atomicMin(first_flag, (int)(p < 2147483646 ? p + 1 : 2147483646));
            }
        }
    }
}

template <int NS, int PTS = J2<NS>::PTS, bool SPEC = false>
static void jac2_launch(const WebDev &w, int nsd, const double *c, const double *rewt, double cj, double *bd,
                        int *ipbd, int *d_first_flag)
{
    static bool attr_set = false;
    const size_t sh = sizeof(double) * (PTS * NS * (NS + 1) + 2 * PTS * NS);
    if (!attr_set) {
        cudaFuncSetAttribute(web_jac2_kernel<NS, PTS, SPEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
        attr_set = true;
    }
    const i64 npts = (i64)w.mx * w.my;
    const int grid = (int)((npts + PTS - 1) / PTS);
    web_jac2_kernel<NS, PTS, SPEC><<<grid, NS * PTS, sh, g_ctx->stream>>>(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag);
    g_ctx->launches++;
}

bool web_jac2(const WebDev &w, int nsd, const double *c, const double *rewt, double cj, double *bd, int *ipbd,
              int *d_first_flag)
{
    switch (w.ns) {
    case 2: jac2_launch<2>(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag); return true;
    case 4: jac2_launch<4>(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag); return true;
    case 6: jac2_launch<6>(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag); return true;
    case 8: jac2_launch<8>(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag); return true;
    case 10: jac2_launch<10>(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag); return true;
    case 12: jac2_launch<12>(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag); return true;
    case 16: jac2_launch<16>(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag); return true;
    case 20:
        switch (g_ctx->debug_variant - 10) {   // 10..15 select this file (web_jac_launch)
        case 1: jac2_launch<20, 16, true>(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag); break;
        case 2: jac2_launch<20, 32, false>(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag); break;
        case 3: jac2_launch<20, 32, true>(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag); break;
        case 4: jac2_launch<20, 8, false>(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag); break;
        case 5: jac2_launch<20, 8, true>(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag); break;
        default: jac2_launch<20, 16, false>(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag); break;
        }
        return true;
    case 24: jac2_launch<24>(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag); return true;
    case 32: jac2_launch<32>(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag); return true;
    default: return false;
    }
}
