// lu.cu -- batched LINPACK dgefa / dgesl for the ns x ns reaction blocks
// (K5/K6 of SURVEY.md 2.3; src/linpack.f90:325-431, 433-520).
//
// Layout is the reference's (src/daskr_rbdpre.f90:237-249): block p at
// bd + p*ns*ns, column-major; pivots 1-based int32 at ipbd + p*ns.
// One sub-warp group of G = 2^ceil(log2 ns) lanes owns one block; lane i holds
// ROW i of the block in registers (a[j] = A(i,j)), so a column is one coalesced
// ns*8-byte segment per load and every elimination / substitution step is a
// width-G shuffle broadcast plus one multiply-add per lane.  Arithmetic is the
// reference's operation for operation (negated multipliers, first-max pivot,
// separate multiply and add via -fmad=false), which is what makes the pivot
// indices bit-identical to the CPU path.
#include "dkb_internal.h"

#include "lu_device.cuh"
#include <algorithm>

// ------------------------------------------------------------------ kernels
template <int NS>
__global__ void __launch_bounds__(256) dgefa_kernel(double *bd, int *ipbd, i64 nblk, int *info, int *first_flag)
{
    constexpr int G = GroupOf<NS>::G;
    const int li = threadIdx.x % G;
    const i64 gpb = blockDim.x / G;
    const i64 ngroups = (i64)gridDim.x * gpb;
    const i64 nloop = (nblk + ngroups - 1) / ngroups;   // uniform trip count: shuffles need whole groups
    for (i64 it = 0; it < nloop; ++it) {
        i64 p = (i64)blockIdx.x * gpb + threadIdx.x / G + it * ngroups;
        const bool live = p < nblk;
        if (!live) p = nblk - 1;
        double *blk = bd + p * (NS * NS);
        double a[NS];
        group_load_block<NS>(blk, li, a);
        int ipv, inf;
        if (NS >= 10) {
            // logical row interchanges: each finished column is stored at its LINPACK position
            // (every lane has its row in registers before the first group-wide reduction, so the
            // in-place stores cannot overtake another lane's loads).  Faster than the exchanging
            // variant from ns ~ 10 up (measured, tools/bench_extra.py); for small blocks the three
            // warp reductions per step cost more than the butterfly they replace.
            inf = group_dgefa_track<NS, false>(a, li, ipv, [&](int col, int pos, double val) {
                if (live && li < NS) blk[col * NS + pos] = val;
            });
        } else {
            inf = group_dgefa<NS>(a, li, ipv);
            if (live && li < NS) {
#pragma unroll
                for (int j = 0; j < NS; ++j) blk[j * NS + li] = a[j];
            }
        }
        if (live && li < NS) {
            ipbd[p * NS + li] = ipv;
            if (li == 0) {
                if (info) info[p] = inf;
                if (inf != 0 && first_flag) atomicMin(first_flag, (int)(p < 2147483646 ? p + 1 : 2147483646));
            }
        }
    }
}

// dgesl on the reference layout, one THREAD per block: the substitution chains of 32 blocks are
// interleaved in a warp and there are no shuffles.  A thread walks its own contiguous block, so a
// warp-wide load touches 32 different lines; each 32-byte sector is consumed by 4 consecutive loads
// of the same thread and stays in L1 in between, which keeps DRAM traffic at one pass over bd.
// Entries [LO, HI) of one block column (NS doubles at `c`, 32-byte aligned when NS % 4 == 0, 16-byte when NS is even)
// with the widest aligned loads: the kernel is bound by L1 tag look-ups (every lane of a warp reads a different line),
// so a 256-bit load moves four entries for the price of one.  Entries outside [LO, HI) of a touched quad are loaded and
// dropped (they belong to the same column).
template <int NS, int LO, int HI>
__device__ __forceinline__ void load_col(const double *__restrict__ c, double (&col)[NS])
{
    if constexpr (NS % 4 == 0) {
#pragma unroll
        for (int q = LO / 4; q < (HI + 3) / 4; ++q) {
            double x0, x1, x2, x3;
            asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x0), "=d"(x1), "=d"(x2), "=d"(x3) : "l"(c + 4 * q));
            col[4 * q] = x0;
            col[4 * q + 1] = x1;
            col[4 * q + 2] = x2;
            col[4 * q + 3] = x3;
        }
    } else if constexpr (NS % 2 == 0) {
#pragma unroll
        for (int q = LO / 2; q < (HI + 1) / 2; ++q) {
            const double2 x = __ldg(reinterpret_cast<const double2 *>(c) + q);
            col[2 * q] = x.x;
            col[2 * q + 1] = x.y;
        }
    } else {
#pragma unroll
        for (int i = LO; i < HI; ++i) col[i] = __ldg(c + i);
    }
}

template <int NS, int K>
__device__ __forceinline__ void tpp_forward(const double *__restrict__ a, const int *__restrict__ pv, double (&b)[NS])
{
    if constexpr (K < NS - 1) {
        const int l = __ldg(pv + K) - 1;
        double col[NS];
        load_col<NS, K + 1, NS>(a + K * NS, col);
        const double bk = b[K];
        double t = bk;
#pragma unroll
        for (int i = K + 1; i < NS; ++i)
            if (i == l) {
                t = b[i];
                b[i] = bk;
            }
        b[K] = t;
#pragma unroll
        for (int i = K + 1; i < NS; ++i) b[i] = b[i] + t * col[i];
        tpp_forward<NS, K + 1>(a, pv, b);
    }
}
template <int NS, int K>
__device__ __forceinline__ void tpp_backward(const double *__restrict__ a, double (&b)[NS])
{
    if constexpr (K >= 0) {
        double col[NS];
        load_col<NS, 0, K + 1>(a + K * NS, col);
        b[K] = b[K] / col[K];
        const double t = -b[K];
#pragma unroll
        for (int i = 0; i < K; ++i) b[i] = b[i] + t * col[i];
        tpp_backward<NS, K - 1>(a, b);
    }
}

template <int NS>
__global__ void __launch_bounds__(128) dgesl_tpp_ref_kernel(const double *__restrict__ bd, const int *__restrict__ ipbd,
                                                            i64 nblk, double *__restrict__ bv)
{
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < nblk; p += stride) {
        const double *a = bd + p * (NS * NS);
        const int *pv = ipbd + p * NS;
        double b[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) b[i] = bv[p * NS + i];
        tpp_forward<NS, 0>(a, pv, b);
        tpp_backward<NS, NS - 1>(a, b);
#pragma unroll
        for (int i = 0; i < NS; ++i) bv[p * NS + i] = b[i];
    }
}

// psol_rbgpre, src/daskr_rbgpre.f90:278-304: every mesh point is solved with the factors of its group.  Thread per
// point; the ngrp blocks (25 x 3.2 KB in the example) stay in L1/L2 and the 32 points of a warp mostly share one.
template <int NS>
__global__ void __launch_bounds__(128) rbg_psol_kernel(const double *__restrict__ bd, const int *__restrict__ ipbd,
                                                       const int *__restrict__ jigx, const int *__restrict__ jigy,
                                                       int mx, int jy0, int ngx, i64 npts, double *__restrict__ bv)
{
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += stride) {
        const int jyl = (int)(p / mx), jx = (int)(p - (i64)jyl * mx);
        const int ig = (__ldg(jigy + jy0 + jyl) - 1) * ngx + (__ldg(jigx + jx) - 1);
        const double *a = bd + (i64)ig * (NS * NS);
        const int *pv = ipbd + (i64)ig * NS;
        double b[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) b[i] = bv[p * NS + i];
#pragma unroll
        for (int k = 0; k < NS - 1; ++k) {
            const int l = __ldg(pv + k) - 1;
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
            for (int i = k + 1; i < NS; ++i) b[i] = b[i] + t * __ldg(a + k * NS + i);
        }
#pragma unroll
        for (int k = NS - 1; k >= 0; --k) {
            b[k] = b[k] / __ldg(a + k * NS + k);
            const double t = -b[k];
#pragma unroll
            for (int i = 0; i < k; ++i) b[i] = b[i] + t * __ldg(a + k * NS + i);
        }
#pragma unroll
        for (int i = 0; i < NS; ++i) bv[p * NS + i] = b[i];
    }
}

__global__ void rbg_add_cj_kernel(double *bd, int ns, int nsd, int ngrp, double cj)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;   // (group, diagonal index)
    if (e >= ngrp * nsd) return;
    const int g = e / nsd, i = e - g * nsd;
    double *d = bd + (i64)g * ns * ns + (i64)i * (ns + 1);
    *d = *d + cj;   // jac_rbgpre:259-262
}

template <int NS>
__global__ void __launch_bounds__(256) dgesl_kernel(const double *__restrict__ bd, const int *__restrict__ ipbd,
                                                    i64 nblk, double *b)
{
    constexpr int G = GroupOf<NS>::G;
    const int li = threadIdx.x % G;
    const i64 gpb = blockDim.x / G;
    const i64 ngroups = (i64)gridDim.x * gpb;
    const i64 nloop = (nblk + ngroups - 1) / ngroups;
    for (i64 it = 0; it < nloop; ++it) {
        i64 p = (i64)blockIdx.x * gpb + threadIdx.x / G + it * ngroups;
        const bool live = p < nblk;
        if (!live) p = nblk - 1;
        double a[NS];
        group_load_block<NS>(bd + p * (NS * NS), li, a);
        int ipv = (li < NS) ? __ldg(ipbd + p * NS + li) : 1;
        double bi = (li < NS) ? b[p * NS + li] : 0.0;
        bi = group_dgesl<NS>(a, ipv, li, bi);
        if (live && li < NS) b[p * NS + li] = bi;
    }
}

// dspigm prologue (:3277-3289): out = (wscale*wt) * P^-1 in, plus sum of squares of out
template <int NS>
__global__ void __launch_bounds__(256) psol_scaled_kernel(const double *__restrict__ bd, const int *__restrict__ ipbd,
                                                          i64 nblk, const double *__restrict__ in,
                                                          const double *__restrict__ wt, double wscale,
                                                          double *out, double *part, unsigned *ticket, double *res)
{
    constexpr int G = GroupOf<NS>::G;
    __shared__ double sh[8];
    __shared__ int is_last;
    const int li = threadIdx.x % G;
    const i64 gpb = blockDim.x / G;
    const i64 ngroups = (i64)gridDim.x * gpb;
    const i64 nloop = (nblk + ngroups - 1) / ngroups;
    double acc = 0.0;
    for (i64 it = 0; it < nloop; ++it) {
        i64 p = (i64)blockIdx.x * gpb + threadIdx.x / G + it * ngroups;
        const bool live = p < nblk;
        if (!live) p = nblk - 1;
        double a[NS];
        group_load_block<NS>(bd + p * (NS * NS), li, a);
        int ipv = (li < NS) ? __ldg(ipbd + p * NS + li) : 1;
        double bi = (li < NS) ? in[p * NS + li] : 0.0;
        bi = group_dgesl<NS>(a, ipv, li, bi);
        if (live && li < NS) {
            double w = wscale * wt[p * NS + li];
            double o = bi * w;
            out[p * NS + li] = o;
            acc = acc + o * o;
        }
    }
    // block reduction + last-block final stage (same scheme as vecops.cu)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) sh[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
        part[(i64)blockIdx.x * 8] = s;
        __threadfence();
        unsigned t = atomicInc(ticket, gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double s = 0.0;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) s += __ldcg(&part[(i64)b * 8]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        __syncthreads();
        if (lane == 0) sh[wid] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += sh[w];
            res[0] = tot;
        }
    }
}

static int lu_grid(i64 nblk, int G)
{
    i64 gpb = 256 / G;
    i64 g = (nblk + gpb - 1) / gpb;
    i64 cap = (i64)g_ctx->nsm * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// every block size the reference accepts up to one warp per block (src/daskr_rbdpre.f90:102-156 takes any ns)
#define NS_DISPATCH(ns, CALL) \
    switch (ns) { \
    case 1: { CALL(1); } break; \
    case 2: { CALL(2); } break; \
    case 3: { CALL(3); } break; \
    case 4: { CALL(4); } break; \
    case 5: { CALL(5); } break; \
    case 6: { CALL(6); } break; \
    case 7: { CALL(7); } break; \
    case 8: { CALL(8); } break; \
    case 9: { CALL(9); } break; \
    case 10: { CALL(10); } break; \
    case 11: { CALL(11); } break; \
    case 12: { CALL(12); } break; \
    case 13: { CALL(13); } break; \
    case 14: { CALL(14); } break; \
    case 15: { CALL(15); } break; \
    case 16: { CALL(16); } break; \
    case 17: { CALL(17); } break; \
    case 18: { CALL(18); } break; \
    case 19: { CALL(19); } break; \
    case 20: { CALL(20); } break; \
    case 21: { CALL(21); } break; \
    case 22: { CALL(22); } break; \
    case 23: { CALL(23); } break; \
    case 24: { CALL(24); } break; \
    case 25: { CALL(25); } break; \
    case 26: { CALL(26); } break; \
    case 27: { CALL(27); } break; \
    case 28: { CALL(28); } break; \
    case 29: { CALL(29); } break; \
    case 30: { CALL(30); } break; \
    case 31: { CALL(31); } break; \
    case 32: { CALL(32); } break; \
    default: return dkb_fail(DKB_E_UNSUPPORTED, "block size ns=%d out of range 1..32", ns); \
    }

int lu_dgefa_batched(double *bd, int *ipbd, int ns, i64 nblk, int *info, int *d_first_flag)
{
    if (nblk <= 0) return 0;
    cudaStream_t st = g_ctx->stream;
#define CALL(N) dgefa_kernel<N><<<lu_grid(nblk, GroupOf<N>::G), 256, 0, st>>>(bd, ipbd, nblk, info, d_first_flag)
    NS_DISPATCH(ns, CALL)
#undef CALL
    COUNT_LAUNCH();
    return 0;
}

int lu_dgesl_batched(const double *bd, const int *ipbd, int ns, i64 nblk, double *b)
{
    if (nblk <= 0) return 0;
    cudaStream_t st = g_ctx->stream;
    // measured (tools/bench_extra.py, 1e6 blocks): one thread per block wins for ns <= 2 and ns >= 20,
    // one sub-warp group per block in between
    static int tpp_min = -1;
    if (tpp_min < 0) {
        const char *e = getenv("DKB_DGESL_TPP_MIN");   // tuning hook: smallest ns (> 2) solved by the thread-per-block kernel
        tpp_min = e ? atoi(e) : 8;
    }
    const bool aligned = (reinterpret_cast<uintptr_t>(bd) & 31u) == 0;   // the thread-per-block kernel uses 128 / 256-bit loads
    if ((ns <= 2 || ns >= tpp_min) && aligned) {
        const int grid = (int)std::min<i64>((nblk + 127) / 128, g_ctx->nsm * 16);
#define CALL(N) dgesl_tpp_ref_kernel<N><<<grid, 128, 0, st>>>(bd, ipbd, nblk, b)
        NS_DISPATCH(ns, CALL)
#undef CALL
    } else {
#define CALL(N) dgesl_kernel<N><<<lu_grid(nblk, GroupOf<N>::G), 256, 0, st>>>(bd, ipbd, nblk, b)
        NS_DISPATCH(ns, CALL)
#undef CALL
    }
    COUNT_LAUNCH();
    return 0;
}

void rbg_add_cj(double *bd, int ns, int nsd, int ngrp, double cj)
{
    const int n = ngrp * nsd;
    if (n <= 0) return;
    rbg_add_cj_kernel<<<(n + 127) / 128, 128, 0, g_ctx->stream>>>(bd, ns, nsd, ngrp, cj);
    COUNT_LAUNCH();
}

int rbg_psol(const double *bd, const int *ipbd, double *b)
{
    DkbCtx *c = g_ctx;
    const i64 npts = c->npts;
    const int grid = (int)std::min<i64>((npts + 127) / 128, g_ctx->nsm * 16);
#define CALL(N) rbg_psol_kernel<N><<<grid, 128, 0, c->stream>>>(bd, ipbd, c->d_jigx, c->d_jigy, c->mx, c->jy0, c->ngx, npts, b)
    NS_DISPATCH(c->ns, CALL)
#undef CALL
    COUNT_LAUNCH();
    return 0;
}

int lu_psol_scaled(const double *bd, const int *ipbd, int ns, i64 nblk, const double *in, const double *wt,
                   double wscale, double *out, int slot)
{
    DkbCtx *c = g_ctx;
    if (nblk <= 0) return 0;
    cudaStream_t st = c->stream;
#define CALL(N)                                                                                                   \
    psol_scaled_kernel<N><<<lu_grid(nblk, GroupOf<N>::G), 256, 0, st>>>(bd, ipbd, nblk, in, wt, wscale, out,    \
                                                                        c->d_part, c->d_ticket, c->d_scal + slot)
    NS_DISPATCH(ns, CALL)
#undef CALL
    COUNT_LAUNCH();
    allreduce_sum(slot, 1);
    return 0;
}
