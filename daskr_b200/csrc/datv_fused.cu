// datv_fused.cu -- K3: the whole Jacobian-vector product of GMRES (datv, src/daskr_new.f90:3429-3520)
// for the food-web problem with the reaction preconditioner, in ONE kernel:
//
//   vnext = wght * P^-1 ( G(t, y + v/wght, ydot + cj*v/wght) - savr )
//
// i.e. datv:3493-3498 (perturbed state), res/f/rates (example_web.f90:190-299), datv:3509
// (difference quotient), psol_rbdpre/dgesl (daskr_rbdpre.f90:253-279) and datv:3518 (rescale),
// without materialising z, ydottemp or vtemp.  Algorithmic traffic per unknown: read v, wght, y,
// ydot, savr, the LU block (ns doubles) and its pivots (4 B), write vnext = 8*(6+ns)+4 bytes.
//
// CTA = 128 threads = a tile of 32 x 4 mesh points; warp w owns row w of the tile, i.e. 32
// consecutive mesh points = exactly one interleaved LU tile (psol_tpp.cu), so every LU load is a
// coalesced 256-byte segment.
//   phase 1 (thread per unknown, coalesced): z = y + v/wght for the tile and its stencil halo
//            -> shared Z[point][ns+1]; ydottemp for the differential species -> shared YDT.
//   phase 2 (thread per mesh point): interaction sum with acoef from the constant bank (uniform
//            operand, no loads), 5-point stencil from Z, residual, minus savr -> b[ns] in registers;
//            then dgesl on b with the LU block streamed from HBM (the memory-bound part: CTAs in
//            phase 1 overlap with CTAs streaming LU).
//   phase 3 (thread per unknown, coalesced): vnext = b * wght through shared memory.
// The arithmetic is the reference's, operation for operation (-fmad=false), so the result is
// bit-identical to the two-kernel path and to the CPU path.
#include "web_dev.cuh"

#define DF_TX 32

__constant__ double c_acoef[DKB_NS_MAX * DKB_NS_MAX];   // column-major (k,i) at [i*ns+k]
__constant__ double c_bcoef[DKB_NS_MAX], c_cox[DKB_NS_MAX], c_coy[DKB_NS_MAX];

void datv_fused_upload_tables(const WebPar &p)
{
    cudaMemcpyToSymbol(c_acoef, p.acoef, sizeof(double) * p.ns * p.ns);
    cudaMemcpyToSymbol(c_bcoef, p.bcoef, sizeof(double) * p.ns);
    cudaMemcpyToSymbol(c_cox, p.cox, sizeof(double) * p.ns);
    cudaMemcpyToSymbol(c_coy, p.coy, sizeof(double) * p.ns);
}

// MODE 0: the whole of datv (perturbed residual, - savr, block solve, * wght).
// MODE 1: the residual half only, vnext = G(y + v/wght, ydot + cj*v/wght) - savr   (datv with jpre = 3: the
//         Gauss-Seidel factor runs between the residual and the block solve).
// MODE 2: the plain residual, vnext = G(y, ydot)                                   (res, example_web.f90:190-224).
template <int NS, int DF_TY, int MINB, int MODE = 0>
__global__ void __launch_bounds__(DF_TX * DF_TY, MINB)
web_datv_single_kernel(WebDev w, const double *__restrict__ v, double vscale, const double *__restrict__ wt, double wscale,
                       const double *__restrict__ y, const double *__restrict__ ydot,
                       const double *__restrict__ savr, double cj, const double *__restrict__ lut,
                       const int *__restrict__ pt, double *__restrict__ vnext)
{
    constexpr int DF_THREADS = DF_TX * DF_TY;
    constexpr int ZX = DF_TX + 2, ZY = DF_TY + 2;
    constexpr int PZ = NS + 1;                 // pitch of a point's species in Z (odd: conflict-free)
    constexpr int NPH = NS / 2;                // differential species (prey) come first
    constexpr int PY = NPH | 1;                // odd pitch for YDT
    extern __shared__ double sm[];
    double *Z = sm;                            // [ZY*ZX][PZ]; reused as the output staging buffer
    double *YDT = sm + ZY * ZX * PZ;           // [DF_THREADS][PY]
    const int tid = threadIdx.x;
    const int tiles_x = w.mx / DF_TX;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int x0 = tx * DF_TX, y0 = ty * DF_TY;
    const i64 mxns = (i64)w.mx * NS;

    // ---- phase 1: perturbed state for tile + halo (datv:3493-3498)
    // Loads are made unconditional (out-of-range cells read a safe in-tile address and are discarded)
    // and issued a halo row at a time, so each thread has 4*IT independent loads in flight instead
    // of a load -> divide -> store chain per element.
    // Row by row: a halo row of the tile is ZX*NS contiguous unknowns in memory, so the flat position e in the row
    // gives the address by one addition and the place in Z by e + e/NS (the earlier whole-tile flat index cost ~40
    // integer instructions per element, 40 % of the kernel's instructions in the ncu source view).
    constexpr int ROWE = ZX * NS;
    constexpr int IT = (ROWE + DF_THREADS - 1) / DF_THREADS;
    const i64 gsafe = (i64)y0 * mxns + (i64)x0 * NS;   // first unknown of the tile: always in range
#pragma unroll 1
    for (int zr = 0; zr < ZY; ++zr) {
        const int gy = y0 + zr - 1;
        const bool rowok = (gy >= 0 && gy < w.my) || (gy == -1 && w.has_lo) || (gy == w.my && w.has_hi);
        const bool rowyd = zr >= 1 && zr <= DF_TY && gy < w.my;
        const i64 rbase = (i64)gy * mxns + (i64)(x0 - 1) * NS;      // e = 0 is species 0 of column x0 - 1
        double rwt[IT], rv[IT], ry[IT], ryd[IT];
        int zo[IT], yo[IT];
#pragma unroll
        for (int u = 0; u < IT; ++u) {
            const int e = tid + u * DF_THREADS;
            const int zx = e / NS, i = e - zx * NS;
            const int gx = x0 + zx - 1;
            const bool ok = (e < ROWE) && rowok && gx >= 0 && gx < w.mx;
            const bool okyd = ok && rowyd && i < NPH && zx >= 1 && zx <= DF_TX;
            const i64 g = ok ? rbase + e : gsafe;
            zo[u] = ok ? zr * (ZX * PZ) + e + zx : -1;
            yo[u] = okyd ? ((zr - 1) * DF_TX + (zx - 1)) * PY + i : -1;
            rwt[u] = (MODE == 2) ? 1.0 : __ldg(wt + g);
            rv[u] = (MODE == 2) ? 0.0 : __ldg(v + g);
            ry[u] = __ldg(y + g);
            ryd[u] = __ldg(ydot + (okyd ? g : gsafe));
        }
#pragma unroll
        for (int u = 0; u < IT; ++u) {
            if (zo[u] >= 0) {
                if (MODE == 2) {
                    Z[zo[u]] = ry[u];
                    if (yo[u] >= 0) YDT[yo[u]] = ryd[u];
                } else {
                    const double wgt = wscale * rwt[u];
                    const double vtmp = (vscale * rv[u]) / wgt;    // vtemp = v/wght, v = vscale * (unnormalised basis vector)
                    Z[zo[u]] = ry[u] + vtmp;                       // z = y + vtemp
                    if (yo[u] >= 0) YDT[yo[u]] = ryd[u] + vtmp * cj;   // ydottemp = ydot + vtemp*cj
                }
            }
        }
    }
    __syncthreads();

    // ---- phase 2: one mesh point per thread
    const int wrow = tid >> 5, lane = tid & 31;
    const int gy = y0 + wrow, gx = x0 + lane;
    const bool live = gy < w.my;
    const i64 p = (i64)gy * w.mx + gx;
    double b[NS];
    if (live) {
        const int jyg = w.jy0 + gy;
        const int pz = (wrow + 1) * ZX + (lane + 1);
        // mirrored Neumann neighbours (f:243-254)
        const int pyl = (jyg == 0) ? pz + ZX : pz - ZX;
        const int pyu = (jyg == w.my_g - 1) ? pz - ZX : pz + ZX;
        const int pxl = (gx == 0) ? pz + 1 : pz - 1;
        const int pxu = (gx == w.mx - 1) ? pz - 1 : pz + 1;
        double zc[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) zc[i] = Z[pz * PZ + i];
        // rates:290-293: rxy(:) = sum_i c(i)*acoef(:,i), accumulated in the order i = 1..ns
#pragma unroll
        for (int k = 0; k < NS; ++k) b[k] = 0.0;
#pragma unroll
        for (int i = 0; i < NS; ++i)
#pragma unroll
            for (int k = 0; k < NS; ++k) b[k] = b[k] + zc[i] * c_acoef[i * NS + k];
        const double yy = jyg * w.dy, xx = gx * w.dx;
        const double fac = 1.0 + w.alpha * xx * yy + w.beta * __ldg(w.sx + gx) * __ldg(w.sy + jyg);
        const i64 g0 = p * NS;
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            const double ci = zc[k];
            const double R = ci * (c_bcoef[k] * fac + b[k]);                       // rates:295-297
            const double dcyli = ci - Z[pyl * PZ + k], dcyui = Z[pyu * PZ + k] - ci;
            const double dcxli = ci - Z[pxl * PZ + k], dcxui = Z[pxu * PZ + k] - ci;
            const double f = c_coy[k] * (dcyui - dcyli) + c_cox[k] * (dcxui - dcxli) + R;   // f:264
            const double d = (k >= NPH) ? -f : (YDT[tid * PY + (k < NPH ? k : 0)] - f);      // res:215-219
            // z = vtemp - savr (datv:3509).  MODE 1 subtracts in the store phase, where savr is read coalesced (here every
            // lane would walk its own 160-byte segment: 20 loads of 32 sectors each)
            b[k] = (MODE == 0) ? d - __ldg(savr + g0 + k) : d;
        }
    }
    __syncthreads();   // Z is free from here on (reused as the output staging buffer)

    if (live && MODE != 0) {
#pragma unroll
        for (int i = 0; i < NS; ++i) Z[tid * PZ + i] = b[i];
    }
    if (live && MODE == 0) {
        // psol_rbdpre: dgesl with the LU block streamed from the interleaved layout
        const double *a = lut + ((p >> 5) * (NS * NS)) * 32 + lane;
        const int *pv = pt + ((p >> 5) * NS) * 32 + lane;
#pragma unroll
        for (int k = 0; k < NS - 1; ++k) {
            const int l = __ldg(pv + k * 32) - 1;
            double col[NS];
#pragma unroll
            for (int i = k + 1; i < NS; ++i) col[i] = __ldg(a + (k * NS + i) * 32);
            const double bk = b[k];
            double t = bk;
#pragma unroll
            for (int i = k + 1; i < NS; ++i)
                if (i == l) {
                    t = b[i];
                    b[i] = bk;
                }
            b[k] = t;
#pragma unroll
            for (int i = k + 1; i < NS; ++i) b[i] = b[i] + t * col[i];
        }
#pragma unroll
        for (int k = NS - 1; k >= 0; --k) {
            double col[NS];
#pragma unroll
            for (int i = 0; i <= k; ++i) col[i] = __ldg(a + (k * NS + i) * 32);
            b[k] = b[k] / col[k];
            const double t = -b[k];
#pragma unroll
            for (int i = 0; i < k; ++i) b[i] = b[i] + t * col[i];
        }
#pragma unroll
        for (int i = 0; i < NS; ++i) Z[tid * PZ + i] = b[i];
    }
    __syncthreads();

    // ---- phase 3: z = z*wght (datv:3518), coalesced rows of 32*ns doubles; NS iterations per
    // thread, weights (MODE 0) / savr (MODE 1) loaded up front (unconditionally, clamped) so the loads overlap
    {
        double rw[NS];
        i64 go[NS];
#pragma unroll
        for (int m = 0; m < NS; ++m) {
            const int idx = tid + m * DF_THREADS;
            const int r = idx / (DF_TX * NS), rem = idx - r * (DF_TX * NS);
            const int oy = y0 + r;
            go[m] = (oy < w.my) ? ((i64)oy * mxns + (i64)x0 * NS + rem) : -1;
            rw[m] = (MODE == 2) ? 0.0 : __ldg((MODE == 0 ? wt : savr) + (go[m] >= 0 ? go[m] : gsafe));
        }
#pragma unroll
        for (int m = 0; m < NS; ++m) {
            const int idx = tid + m * DF_THREADS;
            const int r = idx / (DF_TX * NS), rem = idx - r * (DF_TX * NS);
            const int q = r * DF_TX + rem / NS;
            const double zv = Z[q * PZ + (rem % NS)];
            if (go[m] >= 0) vnext[go[m]] = (MODE == 0) ? zv * (wscale * rw[m]) : (MODE == 1) ? zv - rw[m] : zv;
        }
    }
}

template <int NS, int DF_TY, int MINB, int MODE = 0>
static bool launch_single(const WebDev &w, const double *v, double vscale, const double *wt, double wscale, const double *y,
                          const double *ydot, const double *savr, double cj, const double *lut, const int *pt,
                          double *vnext)
{
    static bool attr_set = false;
    constexpr int DF_THREADS = DF_TX * DF_TY;
    constexpr int PZ = NS + 1, PY = (NS / 2) | 1;
    const size_t sh = sizeof(double) * ((DF_TX + 2) * (DF_TY + 2) * PZ + DF_THREADS * PY);
    if (!attr_set) {
        cudaFuncSetAttribute(web_datv_single_kernel<NS, DF_TY, MINB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
        attr_set = true;
    }
    const int tiles = (w.mx / DF_TX) * ((w.my + DF_TY - 1) / DF_TY);
    web_datv_single_kernel<NS, DF_TY, MINB, MODE><<<tiles, DF_THREADS, sh, g_ctx->stream>>>(w, v, vscale, wt, wscale, y, ydot, savr, cj, lut, pt, vnext);
    g_ctx->launches++;
    return true;
}

bool web_datv_single(const WebDev &w, const double *v, double vscale, const double *wt, double wscale, const double *y,
                     const double *ydot, const double *savr, double cj, const double *lut, const int *pt,
                     double *vnext)
{
    // a warp must coincide with one 32-point LU tile: rows must be a whole number of tiles
    if (w.mx % DF_TX != 0 || w.np * 2 != w.ns) return false;
    ProfScope ps(PROF_DATV1);
    switch (w.ns) {
    case 2: return launch_single<2, 4, 4>(w, v, vscale, wt, wscale, y, ydot, savr, cj, lut, pt, vnext);
    case 4: return launch_single<4, 4, 4>(w, v, vscale, wt, wscale, y, ydot, savr, cj, lut, pt, vnext);
    case 8: return launch_single<8, 4, 4>(w, v, vscale, wt, wscale, y, ydot, savr, cj, lut, pt, vnext);
    case 10: return launch_single<10, 4, 4>(w, v, vscale, wt, wscale, y, ydot, savr, cj, lut, pt, vnext);
    case 12: return launch_single<12, 4, 4>(w, v, vscale, wt, wscale, y, ydot, savr, cj, lut, pt, vnext);
    case 16: return launch_single<16, 4, 4>(w, v, vscale, wt, wscale, y, ydot, savr, cj, lut, pt, vnext);
    case 20:
        switch (g_ctx->debug_variant) {
        case 1: return launch_single<20, 4, 3>(w, v, vscale, wt, wscale, y, ydot, savr, cj, lut, pt, vnext);
        case 2: return launch_single<20, 8, 2>(w, v, vscale, wt, wscale, y, ydot, savr, cj, lut, pt, vnext);
        default: return launch_single<20, 4, 4>(w, v, vscale, wt, wscale, y, ydot, savr, cj, lut, pt, vnext);   // 128 regs, 16 warps/SM: fastest measured
        }
    default: return false;
    }
}

// The residual alone with the same staging: mode 1 = datv's perturbed residual minus savr, mode 2 = plain res.
// Same eligibility as web_datv_single; the caller falls back to web_restile_kernel otherwise.
bool web_resid_single(const WebDev &w, int mode, const double *v, double vscale, const double *wt, double wscale, const double *y,
                      const double *ydot, const double *savr, double cj, double *out)
{
    if (w.mx % DF_TX != 0 || w.np * 2 != w.ns) return false;
    if (mode == 1 && savr == nullptr) return false;   // mode 1 subtracts savr in its store phase
#define RS(N) case N: return mode == 1 ? launch_single<N, 4, 4, 1>(w, v, vscale, wt, wscale, y, ydot, savr, cj, nullptr, nullptr, out) \
                                       : launch_single<N, 4, 4, 2>(w, v, vscale, wt, wscale, y, ydot, savr, cj, nullptr, nullptr, out);
    if (w.ns == 20 && mode == 1) {   // taller tiles re-read fewer halo rows (6/4, 10/8, 18/16 of the tile from L2)
        switch (g_ctx->debug_variant) {
        case 21: return launch_single<20, 8, 2, 1>(w, v, vscale, wt, wscale, y, ydot, savr, cj, nullptr, nullptr, out);
        case 22: return launch_single<20, 16, 1, 1>(w, v, vscale, wt, wscale, y, ydot, savr, cj, nullptr, nullptr, out);
        case 23: return launch_single<20, 8, 3, 1>(w, v, vscale, wt, wscale, y, ydot, savr, cj, nullptr, nullptr, out);
        default: break;
        }
    }
    switch (w.ns) {
        RS(2) RS(4) RS(8) RS(10) RS(12) RS(16) RS(20)
    default: return false;
    }
#undef RS
}
