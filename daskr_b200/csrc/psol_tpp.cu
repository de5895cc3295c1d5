// psol_tpp.cu -- psol_rbdpre (src/daskr_rbdpre.f90:253-279: one LINPACK dgesl per mesh point) as a
// thread-per-point kernel over a tile-interleaved copy of the LU blocks.
//
// Why not the reference's block-contiguous layout: with one (sub-)warp per ns x ns block the
// forward/back substitution is ~2*ns dependent shuffle+divide steps with a single block in flight
// per warp -- latency-bound at 14% of HBM peak (profiles/r01).  Here each THREAD owns one block and
// a warp works on 32 blocks at once, so the dependent chain is 32-way interleaved and there are no
// shuffles at all.  For that to stream at HBM speed the blocks are stored interleaved in groups of
// 32 mesh points ("tiles"):
//     element (i,j) of point p  ->  lut[ ((p/32)*ns*ns + j*ns + i)*32 + p%32 ]
//     pivot k of point p        ->  pt [ ((p/32)*ns    + k     )*32 + p%32 ]
// so that the 32 threads of a warp read 256 contiguous bytes for every matrix element, each element
// exactly once per solve, in the order the substitution consumes them.  The right-hand sides keep
// DASKR's species-fastest layout and are transposed through shared memory (row pitch ns+1 doubles:
// conflict-free) with coalesced global loads/stores on both sides.
//
// The arithmetic per block is dgesl's, statement for statement (pivot swap, b(k+1:n) += t*a(k+1:n,k),
// b(k) /= a(k,k), b(1:k-1) -= b(k)*a(1:k-1,k)), so results are bit-identical to the sub-warp kernels
// and to the CPU path.
#include "dkb_internal.h"
#include <algorithm>

#define TPP_THREADS 128

template <int NS>
__global__ void __launch_bounds__(TPP_THREADS, 4)
tpp_psol_kernel(const double *__restrict__ lut, const int *__restrict__ pt, i64 npts,
                const double *__restrict__ in, const double *__restrict__ savr, const double *__restrict__ wt,
                double wscale, double *__restrict__ out, int do_red, double *part, unsigned *ticket, double *res)
{
    constexpr int PITCH = NS + 1;
    __shared__ double sb[TPP_THREADS * PITCH];
    __shared__ double sred[TPP_THREADS / 32];
    __shared__ int is_last;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const i64 ntile = (npts + TPP_THREADS - 1) / TPP_THREADS;
    double acc = 0.0;

    for (i64 tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
        const i64 p0 = tile * TPP_THREADS;
        const i64 g0 = p0 * NS;
        const i64 nval = min((i64)TPP_THREADS, npts - p0) * NS;
        // ---- phase 1: coalesced load of the tile's right-hand sides, (vtemp - savr) of datv:3509.
        // Unconditional (clamped) loads, all NS per thread issued before the first use.
        {
            double rin[NS], rsv[NS];
#pragma unroll
            for (int m = 0; m < NS; ++m) {
                const int idx = tid + m * TPP_THREADS;
                const i64 g = g0 + (idx < nval ? idx : 0);
                rin[m] = in[g];
                rsv[m] = savr ? savr[g] : 0.0;
            }
#pragma unroll
            for (int m = 0; m < NS; ++m) {
                const int idx = tid + m * TPP_THREADS;
                double x = 0.0;
                if (idx < nval) x = savr ? (rin[m] - rsv[m]) : rin[m];
                sb[(idx / NS) * PITCH + (idx % NS)] = x;
            }
        }
        __syncthreads();

        // ---- phase 2: one dgesl per thread, LU streamed from the interleaved layout
        const i64 p = p0 + tid;
        if (p < npts) {
            const i64 t32 = p >> 5;
            const double *a = lut + (t32 * (NS * NS)) * 32 + lane;   // element e at a[e*32]
            const int *pv = pt + (t32 * NS) * 32 + lane;
            double b[NS];
#pragma unroll
            for (int i = 0; i < NS; ++i) b[i] = sb[tid * PITCH + i];
            // forward elimination
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
            // back substitution
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
            for (int i = 0; i < NS; ++i) sb[tid * PITCH + i] = b[i];
        }
        __syncthreads();

        // ---- phase 3: coalesced store, z = z*wght (datv:3518 / dspigm:3281) and the norm partial
        {
            double rw[NS];
#pragma unroll
            for (int m = 0; m < NS; ++m) {
                const int idx = tid + m * TPP_THREADS;
                rw[m] = wt ? wt[g0 + (idx < nval ? idx : 0)] : 1.0;
            }
#pragma unroll
            for (int m = 0; m < NS; ++m) {
                const int idx = tid + m * TPP_THREADS;
                if (idx < nval) {
                    double x = sb[(idx / NS) * PITCH + (idx % NS)];
                    if (wt) x = x * (wscale * rw[m]);
                    out[g0 + idx] = x;
                    acc = acc + x * x;
                }
            }
        }
        __syncthreads();
    }

    if (!do_red) return;
    // block reduction + last-block-done final stage (fixed order: reproducible)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) sred[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < TPP_THREADS / 32; ++w) s += sred[w];
        part[(i64)blockIdx.x * 8] = s;
        __threadfence();
        unsigned t = atomicInc(ticket, gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        double s = 0.0;
        for (int b = tid; b < (int)gridDim.x; b += TPP_THREADS) s += __ldcg(&part[(i64)b * 8]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        __syncthreads();
        if (lane == 0) sred[tid >> 5] = s;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int w = 0; w < TPP_THREADS / 32; ++w) tot += sred[w];
            res[0] = tot;
        }
    }
}

// reference layout (bd, ipbd) <-> interleaved layout; used by the parity tests and by callers that
// want jac_rbdpre's bd/ipbd exactly as the reference leaves them
template <int DIR>
__global__ void __launch_bounds__(256) tpp_relayout_kernel(int ns, i64 npts, double *ref_bd, int *ref_ip,
                                                           double *lut, int *pt)
{
    const i64 nsq = (i64)ns * ns;
    const i64 total = npts * nsq;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += stride) {
        const i64 p = g / nsq;
        const int e = (int)(g - p * nsq);
        const i64 ti = ((p >> 5) * nsq + e) * 32 + (p & 31);
        if (DIR == 0) lut[ti] = ref_bd[g]; else ref_bd[g] = lut[ti];
        if (e < ns) {
            const i64 tp = ((p >> 5) * ns + e) * 32 + (p & 31);
            if (DIR == 0) pt[tp] = ref_ip[p * ns + e]; else ref_ip[p * ns + e] = pt[tp];
        }
    }
}

#define TPP_DISPATCH(ns, CALL)                                                                       \
    switch (ns) { \
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
    default: return dkb_fail(DKB_E_UNSUPPORTED, "block size ns=%d out of range 2..32", ns); \
    }

// x = in - savr (savr may be null); solve; out = x * (wscale*wt) (wt may be null); slot >= 0: sum out^2
int tpp_psol(const double *lut, const int *pt, int ns, i64 npts, const double *in, const double *savr,
             const double *wt, double wscale, double *out, int slot)
{
    DkbCtx *c = g_ctx;
    if (npts <= 0) return 0;
    i64 ntile = (npts + TPP_THREADS - 1) / TPP_THREADS;
    int grid = (int)std::min<i64>(ntile, g_ctx->nsm * 4);
    double *res = slot >= 0 ? c->d_scal + slot : nullptr;
    {
        ProfScope ps(PROF_TPP);   // events around this one launch: the roofline kernel of bench.py
#define CALL(N)                                                                                              \
    tpp_psol_kernel<N><<<grid, TPP_THREADS, 0, c->stream>>>(lut, pt, npts, in, savr, wt, wscale, out,        \
                                                            slot >= 0 ? 1 : 0, c->d_part, c->d_ticket, res)
        TPP_DISPATCH(ns, CALL)
#undef CALL
    }
    COUNT_LAUNCH();
    if (slot >= 0) allreduce_sum(slot, 1);
    return 0;
}

i64 tpp_lut_doubles(int ns, i64 npts) { return ((npts + 31) / 32) * 32 * (i64)ns * ns; }
i64 tpp_pt_ints(int ns, i64 npts) { return ((npts + 31) / 32) * 32 * (i64)ns; }

int tpp_from_reference(int ns, i64 npts, const double *bd, const int *ipbd, double *lut, int *pt)
{
    if (npts <= 0) return 0;
    i64 total = npts * ns * ns;
    int grid = (int)std::min<i64>((total + 255) / 256, g_ctx->nsm * 8);
    tpp_relayout_kernel<0><<<grid, 256, 0, g_ctx->stream>>>(ns, npts, const_cast<double *>(bd), const_cast<int *>(ipbd), lut, pt);
    COUNT_LAUNCH();
    return 0;
}
int tpp_to_reference(int ns, i64 npts, const double *lut, const int *pt, double *bd, int *ipbd)
{
    if (npts <= 0) return 0;
    i64 total = npts * ns * ns;
    int grid = (int)std::min<i64>((total + 255) / 256, g_ctx->nsm * 8);
    tpp_relayout_kernel<1><<<grid, 256, 0, g_ctx->stream>>>(ns, npts, bd, ipbd, const_cast<double *>(lut), const_cast<int *>(pt));
    COUNT_LAUNCH();
    return 0;
}
