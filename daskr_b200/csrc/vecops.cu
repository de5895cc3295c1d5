// vecops.cu -- length-NEQ vector kernels of the Krylov path (K1, K2, K7, K8 of SURVEY.md 2.3).
//
// Everything here is HBM-bound FP64 streaming: coalesced grid-stride loops at
// full occupancy, warp-shuffle + shared-memory block reductions, and a
// last-block-done final stage so that every dot product / norm is produced by
// ONE launch, stays in device memory (d_scal slots) and is bit-reproducible
// for a given n (fixed grid, fixed summation order, no floating-point atomics).
// Compiled with -fmad=false: the reference arithmetic is separate multiply and
// add, so elementwise results match the CPU path bit for bit.
#include "dkb_internal.h"
#include <cfloat>

// ------------------------------------------------------------------ launch geometry
static inline int map_grid(i64 n, int block)
{
    i64 g = (n + block - 1) / block;
    i64 cap = (i64)g_ctx->nsm * 8;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}
static inline int red_grid(i64 n)
{
    i64 g = (n + DKB_RED_BLOCK * 4 - 1) / (DKB_RED_BLOCK * 4);
    i64 cap = std::min<i64>((i64)g_ctx->nsm * 4, DKB_RED_MAXGRID);   // 4 x 512 threads per SM = full occupancy
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// ------------------------------------------------------------------ generic elementwise map
template <class F>
__global__ void __launch_bounds__(256) map_kernel(F f, i64 n)
{
    const i64 stride = (i64)gridDim.x * blockDim.x;
#pragma unroll 4
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}
template <class F>
static void launch_map(F f, i64 n)
{
    if (n <= 0) return;
    map_kernel<F><<<map_grid(n, 256), 256, 0, g_ctx->stream>>>(f, n);
    COUNT_LAUNCH();
}

struct FCopy { const double *x; double *y; __device__ void operator()(i64 i) const { y[i] = x[i]; } };
struct FZero { double *y; __device__ void operator()(i64 i) const { y[i] = 0.0; } };
struct FScale { double a; double *x; __device__ void operator()(i64 i) const { x[i] = a * x[i]; } };
struct FScaleTo { double a; const double *x; double *y; __device__ void operator()(i64 i) const { y[i] = a * x[i]; } };
struct FAxpy { double a; const double *x; double *y; __device__ void operator()(i64 i) const { y[i] = y[i] + a * x[i]; } };
struct FLin2 { double a, b; const double *x, *y; double *z; __device__ void operator()(i64 i) const { z[i] = a * x[i] + b * y[i]; } };
struct FMul { const double *x, *w; double *y; __device__ void operator()(i64 i) const { y[i] = x[i] * w[i]; } };
struct FSub { const double *x, *y; double *z; __device__ void operator()(i64 i) const { z[i] = x[i] - y[i]; } };

struct FAxpyS { double a, s; const double *x; double *y; __device__ void operator()(i64 i) const { y[i] = y[i] + a * (s * x[i]); } };
struct FLin2S { double a, b, s; const double *x, *y; double *z; __device__ void operator()(i64 i) const { z[i] = a * x[i] + b * (s * y[i]); } };
void v_axpy_s(i64 n, double a, const double *x, double s, double *y) { launch_map(FAxpyS{a, s, x, y}, n); }
void v_lin2_s(i64 n, double a, const double *x, double b, const double *y, double s, double *z) { launch_map(FLin2S{a, b, s, x, y, z}, n); }
void v_copy(i64 n, const double *x, double *y) { if (x != y) launch_map(FCopy{x, y}, n); }
void v_zero(i64 n, double *y) { launch_map(FZero{y}, n); }
void v_scale(i64 n, double a, double *x) { launch_map(FScale{a, x}, n); }
void v_scale_to(i64 n, double a, const double *x, double *y) { launch_map(FScaleTo{a, x, y}, n); }
void v_axpy(i64 n, double a, const double *x, double *y) { launch_map(FAxpy{a, x, y}, n); }
void v_lin2(i64 n, double a, const double *x, double b, const double *y, double *z) { launch_map(FLin2{a, b, x, y, z}, n); }
void v_mul(i64 n, const double *x, const double *w, double *y) { launch_map(FMul{x, w, y}, n); }
void v_sub(i64 n, const double *x, const double *y, double *z) { launch_map(FSub{x, y, z}, n); }

// predictor / interpolation: yout = p1; dout = 0; for j=2..kp1: yout += c_j p_j, dout += d_j p_j
// (dnedk:2808-2813 with c_j = 1, d_j = gama(j); ddatrp:812-824 with its c, d recurrences)
struct FPredict {
    int kp1;
    const double *p[6];
    double c[6], d[6];
    double *yout, *dout;
    __device__ void operator()(i64 i) const
    {
        double yy = p[0][i], dd = 0.0;
#pragma unroll
        for (int j = 1; j < 6; ++j)
            if (j < kp1) {
                double pj = p[j][i];
                yy = yy + c[j] * pj;
                dd = dd + d[j] * pj;
            }
        yout[i] = yy;
        dout[i] = dd;
    }
};
void v_predict(i64 n, int kp1, double *const *phi, const double *c, const double *d, double *yout, double *dout)
{
    FPredict f;
    f.kp1 = kp1;
    for (int j = 0; j < 6; ++j) {
        f.p[j] = phi[j < kp1 ? j : 0];
        f.c[j] = j < kp1 ? c[j] : 0.0;
        f.d[j] = j < kp1 ? d[j] : 0.0;
    }
    f.yout = yout;
    f.dout = dout;
    launch_map(f, n);
}

// ddstp:458-466: if (write_kp2) phi(kp2) = e; phi(kp1) += e; phi(j) += phi(j+1), j = kp1-1..1
struct FPhiUpdate {
    int kp1, write_kp2;
    double *p[7];
    const double *e;
    __device__ void operator()(i64 i) const
    {
        double ee = e[i];
        if (write_kp2) p[kp1][i] = ee;
        double run = p[kp1 - 1][i] + ee;
        p[kp1 - 1][i] = run;
#pragma unroll
        for (int j = 5; j >= 0; --j)
            if (j < kp1 - 1) {
                run = p[j][i] + run;
                p[j][i] = run;
            }
    }
};
void v_phi_update(i64 n, int kp1, int write_kp2, double *const *phi, const double *e)
{
    FPhiUpdate f;
    f.kp1 = kp1;
    f.write_kp2 = write_kp2;
    for (int j = 0; j < 7; ++j) f.p[j] = (j <= kp1 && j < 6) ? phi[j] : phi[0];
    f.e = e;
    launch_map(f, n);
}

// ddawts + dinvwt (+ the vt mask of daskr.f:1755): wt = 1/(rtol*|y| + atol)
struct FEwt {
    int iwt;
    double rtol, atol;
    const double *vrtol, *vatol, *y;
    double *wt, *vt;
    const int *id;
    int *first_bad;
    __device__ void operator()(i64 i) const
    {
        double w = iwt ? (vrtol[i] * fabs(y[i]) + vatol[i]) : (rtol * fabs(y[i]) + atol);
        if (w <= 0.0) {
            int idx = (int)(i + 1);
            atomicMin(first_bad, idx);
        }
        double r = 1.0 / w;
        wt[i] = r;
        if (vt != nullptr && vt != wt) {
            int m = id[i] > 0 ? id[i] : 0;
            vt[i] = (double)m * r;
        }
    }
};
void v_ewt(i64 n, int iwt, double rtol, double atol, const double *d_rtol, const double *d_atol,
           const double *y, double *wt, double *vt, const int *id, int *first_bad)
{
    launch_map(FEwt{iwt, rtol, atol, d_rtol, d_atol, y, wt, vt, id, first_bad}, n);
}

// datv:3493-3498 (unfused form used with user callbacks)
struct FDatvPre {
    const double *v, *wt, *y, *ydot;
    double vscale, wscale, cj;
    double *vtemp, *ydottemp, *z;
    __device__ void operator()(i64 i) const
    {
        double vt = (vscale * v[i]) / (wscale * wt[i]);   // v(:,ll) is kept unnormalised: vscale = 1/snormw (dspigm:3358)
        vtemp[i] = vt;
        ydottemp[i] = ydot[i] + vt * cj;
        z[i] = y[i] + vt;
    }
};
void v_datv_pre(i64 n, const double *v, double vscale, const double *wt, double wscale, const double *y,
                const double *ydot, double cj, double *vtemp, double *ydottemp, double *z)
{
    launch_map(FDatvPre{v, wt, y, ydot, vscale, wscale, cj, vtemp, ydottemp, z}, n);
}

struct FNewton {
    const double *delta;
    double cj;
    double *y, *e, *ydot;
    __device__ void operator()(i64 i) const
    {
        double d = delta[i];
        y[i] = y[i] - d;
        e[i] = e[i] - d;
        ydot[i] = ydot[i] - cj * d;
    }
};
void v_newton_update(i64 n, const double *delta, double cj, double *y, double *e, double *ydot)
{
    launch_map(FNewton{delta, cj, y, e, ydot}, n);
}

// DYYPNW, src/daskr.f:3078-3131
struct FDyypnw {
    const double *y, *yp, *p;
    double cj, rl;
    int icopt;
    const int *id;
    double *ynew, *ypnew;
    __device__ void operator()(i64 i) const
    {
        if (icopt == 1) {
            if (id[i] < 0) {
                ynew[i] = y[i] - rl * p[i];
                ypnew[i] = yp[i];
            } else {
                ynew[i] = y[i];
                ypnew[i] = yp[i] - rl * cj * p[i];
            }
        } else {
            ynew[i] = y[i] - rl * p[i];
            ypnew[i] = yp[i];
        }
    }
};
void v_dyypnw(i64 n, const double *y, const double *yp, double cj, double rl, const double *p,
              int icopt, const int *id, double *ynew, double *ypnew)
{
    launch_map(FDyypnw{y, yp, p, cj, rl, icopt, id, ynew, ypnew}, n);
}

struct FMin0 { const double *y; double *d; __device__ void operator()(i64 i) const { d[i] = fmin(y[i], 0.0); } };
void v_min0(i64 n, const double *y, double *delta) { launch_map(FMin0{y, delta}, n); }

// dspigm:3413-3417 + dslvk:3143-3145: z = 0; z += r_i v_i (daxpy order); z = z/wght; x = x + z
struct FGmresSol {
    int ll, first;
    const double *v[DKB_MAXL_MAX];
    double r[DKB_MAXL_MAX], vs[DKB_MAXL_MAX];   // vs: normalisation factors of the basis vectors (kept unnormalised in memory)
    const double *wt;
    double wscale;
    double *x;
    __device__ void operator()(i64 i) const
    {
        double z = 0.0;
#pragma unroll 1
        for (int k = 0; k < ll; ++k)
            if (r[k] != 0.0) z = z + r[k] * (vs[k] * v[k][i]);   // daxpy skips da == 0
        z = z / (wscale * wt[i]);
        x[i] = first ? (0.0 + z) : (x[i] + z);
    }
};
void v_gmres_solution(i64 n, int ll, double *const *v, const double *vs, const double *r, const double *wt,
                      double wscale, double *x, int first)
{
    FGmresSol f;
    f.ll = ll;
    f.first = first;
    for (int k = 0; k < DKB_MAXL_MAX; ++k) {
        f.v[k] = v[k < ll ? k : 0];
        f.r[k] = k < ll ? r[k] : 0.0;
        f.vs[k] = k < ll ? vs[k] : 0.0;
    }
    f.wt = wt;
    f.wscale = wscale;
    f.x = x;
    launch_map(f, n);
}

// dspigm:3381-3401 (kmp == maxl): dl = v1; dl = s_i*dl + c_i*v_{i+1} (i < maxl);
// dl = s_m*dl + (c_m/snormw)*v_{maxl+1}; dl = scale*dl
struct FGmresDl {
    int maxl;
    const double *v[DKB_MAXL_MAX + 1];
    double s[DKB_MAXL_MAX], c[DKB_MAXL_MAX], vs[DKB_MAXL_MAX];
    double scale;
    double *dl;
    __device__ void operator()(i64 i) const
    {
        double d = vs[0] * v[0][i];
#pragma unroll 1
        for (int k = 1; k < maxl; ++k) d = s[k - 1] * d + c[k - 1] * (vs[k] * v[k][i]);
        d = s[maxl - 1] * d + c[maxl - 1] * v[maxl][i];   // v(:,maxl+1) is not normalised in the reference either (:3395-3398)
        dl[i] = scale * d;
    }
};
void v_gmres_dl(i64 n, int maxl, double *const *v, const double *vs, const double *q, double snormw, double scale, double *dl)
{
    FGmresDl f;
    f.maxl = maxl;
    for (int k = 0; k <= DKB_MAXL_MAX; ++k) f.v[k] = v[k <= maxl ? k : 0];
    for (int k = 0; k < DKB_MAXL_MAX; ++k) f.vs[k] = k < maxl ? vs[k] : 0.0;
    for (int k = 1; k <= maxl; ++k) {
        f.s[k - 1] = q[2 * k - 1];
        f.c[k - 1] = (k < maxl) ? q[2 * k - 2] : q[2 * k - 2] / snormw;
    }
    f.scale = scale;
    f.dl = dl;
    launch_map(f, n);
}

// ------------------------------------------------------------------ reductions
// KS sums followed by KM maxima.  Stage 1: per-thread grid-stride accumulation,
// warp shuffles, shared memory.  Stage 2: the block that takes the last ticket
// adds the per-block partials in a fixed order and writes out[0..KS+KM).
template <int KS, int KM>
__device__ __forceinline__ void block_reduce(double (&acc)[KS + KM], double *sh)
{
    constexpr int K = KS + KM;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double a = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double b = __shfl_xor_sync(0xffffffffu, a, o);
            a = (k < KS) ? (a + b) : fmax(a, b);
        }
        if (lane == 0) sh[wid * K + k] = a;
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double a = (lane < nw) ? sh[lane * K + k] : ((k < KS) ? 0.0 : -DBL_MAX);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double b = __shfl_xor_sync(0xffffffffu, a, o);
                a = (k < KS) ? (a + b) : fmax(a, b);
            }
            acc[k] = a;
        }
    }
    __syncthreads();
}

// Several ranks (pc.nranks > 1): the block that finishes last also performs the all-reduction, through peer memory.
// It stores this rank's K totals into every rank's mailbox (NVLink stores, one thread per destination), publishes them
// with a sequence number, waits for the same sequence number from every source in its own mailbox and combines the
// P contributions in rank order -- the same order on every rank, so all ranks hold bit-identical results (which the
// replicated host control relies on).  Values are double-buffered by the parity of the sequence number: a rank can be
// at most one all-reduction ahead of the slowest one, because every all-reduction needs everybody's contribution.
template <int KS, int KM, class F>
__global__ void __launch_bounds__(DKB_RED_BLOCK) reduce_kernel(F f, i64 n, double *part,
                                                               unsigned *ticket, double *out, const PeerComm pc)
{
    constexpr int K = KS + KM;
    __shared__ double sh[(DKB_RED_BLOCK / 32) * K];
    __shared__ double tot[K];
    __shared__ int is_last;
    double acc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] = (k < KS) ? 0.0 : -DBL_MAX;
    f.init();
    const i64 stride = (i64)gridDim.x * blockDim.x;
#pragma unroll 2
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i, acc);

    block_reduce<KS, KM>(acc, sh);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) part[(i64)blockIdx.x * 8 + k] = acc[k];
        __threadfence();
        unsigned t = atomicInc(ticket, gridDim.x - 1);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = (k < KS) ? 0.0 : -DBL_MAX;
        for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double p = __ldcg(&part[(i64)b * 8 + k]);
                acc[k] = (k < KS) ? (acc[k] + p) : fmax(acc[k], p);
            }
        }
        block_reduce<KS, KM>(acc, sh);
        if (pc.nranks <= 1) {
            if (threadIdx.x == 0) {
#pragma unroll
                for (int k = 0; k < K; ++k) out[k] = acc[k];
            }
            return;
        }
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) tot[k] = acc[k];
        }
        __syncthreads();
        const int par = (int)(pc.seq & 1ull);
        if ((int)threadIdx.x < pc.nranks) {
            const int r = threadIdx.x;
            double *dst = pc.box[r] + DKB_BOX_VAL + (par * DKB_PEER_MAX + pc.rank) * 8;
#pragma unroll
            for (int k = 0; k < K; ++k) dst[k] = tot[k];
            __threadfence_system();
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"((unsigned long long *)(pc.box[r] + DKB_BOX_ARFLAG) + pc.rank), "l"(pc.seq) : "memory");
            // ... and wait for source r's contribution in this rank's own mailbox
            const unsigned long long *fl = (const unsigned long long *)(pc.box[pc.rank] + DKB_BOX_ARFLAG) + r;
            unsigned long long seen;
            do { asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(fl) : "memory"); } while (seen < pc.seq);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const volatile double *src = pc.box[pc.rank] + DKB_BOX_VAL + par * DKB_PEER_MAX * 8;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                double t = (k < KS) ? 0.0 : -DBL_MAX;
                for (int r = 0; r < pc.nranks; ++r) {
                    const double v = src[r * 8 + k];
                    t = (k < KS) ? (t + v) : fmax(t, v);
                }
                out[k] = t;
            }
        }
    }
}

// One launch: local reduction + (several ranks) the all-reduction of the KS sums and KM maxima -- inside the kernel when
// the peer-memory path is up, by NCCL calls behind it otherwise.
template <int KS, int KM, class F>
static void launch_reduce(F f, i64 n, int slot)
{
    DkbCtx *c = g_ctx;
    const PeerComm pc = peer_comm_next();
    reduce_kernel<KS, KM, F><<<red_grid(n), DKB_RED_BLOCK, 0, c->stream>>>(f, n, c->d_part, c->d_ticket,
                                                                            c->d_scal + slot, pc);
    COUNT_LAUNCH();
    if (c->nranks > 1 && pc.nranks <= 1) {
        if (KS > 0) allreduce_sum(slot, KS);
        if (KM > 0) allreduce_max(slot + KS, KM);
    }
}

struct RDot {
    const double *x, *y;
    __device__ void init() {}
    __device__ void operator()(i64 i, double (&a)[1]) const { a[0] = a[0] + x[i] * y[i]; }
};
// The basis vectors v(:,1:ll) stay unnormalised in memory; s1 / si / sn are their normalisation factors 1/||.||
// (dspigm:3295, 3358): s*v[i] is exactly the value the reference's dscal stored.
struct RDot2 {
    const double *vnew, *v1;
    double s1;
    __device__ void init() {}
    __device__ void operator()(i64 i, double (&a)[2]) const
    {
        double w = vnew[i];
        a[0] = a[0] + w * w;
        a[1] = a[1] + (s1 * v1[i]) * w;
    }
};
struct RAxpyDot {
    double *vnew;
    const double *vi, *hptr, *vnext;
    double si, sn;
    double h;
    __device__ void init() { h = *hptr; }
    __device__ void operator()(i64 i, double (&a)[1]) const
    {
        double w = vnew[i] - h * (si * vi[i]);   // dorth:3565-3566 (tem = -hes; vnew += tem*v_i)
        vnew[i] = w;
        a[0] = a[0] + (sn * vnext[i]) * w;
    }
};
struct RAxpyNrm {
    double *vnew;
    const double *vi, *hptr;
    double si;
    double h;
    __device__ void init() { h = *hptr; }
    __device__ void operator()(i64 i, double (&a)[1]) const
    {
        double w = vnew[i] - h * (si * vi[i]);
        vnew[i] = w;
        a[0] = a[0] + w * w;
    }
};
struct RDotS {
    const double *x, *y;
    double sx;
    __device__ void init() {}
    __device__ void operator()(i64 i, double (&a)[1]) const { a[0] = a[0] + (sx * x[i]) * y[i]; }
};
struct RWrms {
    const double *v, *rwt;
    __device__ void init() {}
    __device__ void operator()(i64 i, double (&a)[2]) const
    {
        double p = v[i] * rwt[i];
        a[0] = a[0] + p * p;
        a[1] = fmax(a[1], fabs(p));
    }
};
struct RWrmsScaled {   // the reference's second pass, ddwnrm:861-864
    const double *v, *rwt;
    double vmax;
    __device__ void init() {}
    __device__ void operator()(i64 i, double (&a)[1]) const
    {
        double q = (v[i] * rwt[i]) / vmax;
        a[0] = a[0] + q * q;
    }
};

struct RSumStrided {   // sum of x[i*stride + off]
    const double *x;
    i64 stride, off;
    __device__ void init() {}
    __device__ void operator()(i64 i, double (&a)[1]) const { a[0] = a[0] + x[i * stride + off]; }
};
struct RMaxVal {
    const double *x;
    __device__ void init() {}
    __device__ void operator()(i64 i, double (&a)[1]) const { a[0] = fmax(a[0], x[i]); }
};
void r_sum_strided(i64 npts, i64 stride, i64 off, const double *x, int slot)
{
    launch_reduce<1, 0>(RSumStrided{x, stride, off}, npts, slot);
}
void r_max(i64 n, const double *x, int slot)
{
    launch_reduce<0, 1>(RMaxVal{x}, n, slot);
}
// dcnstr, src/daskr_new.f90:569-661, as one reduction.  The reference scans the unknowns in order, returns at the first
// hard violation (tau = 0.6*tau) and otherwise compares the largest relative change rdymx of the |icnstr| = 2 unknowns
// with rlx.  Both outcomes are order-free: "is there a hard violation" (a max over 0/1 flags) and rdymx (a max).
struct RCnstr {
    const double *y, *ynew;
    const int *icn;
    __device__ void init() {}
    __device__ void operator()(i64 i, double (&a)[2]) const
    {
        const int ic = icn[i];
        if (ic == 0) return;
        const double yn = ynew[i];
        bool hard;
        if (ic == 2 || ic == -2) {
            const double yi = y[i];
            a[0] = fmax(a[0], fabs((yn - yi) / yi));
            hard = (ic == 2) ? (yn <= 0.0) : (yn >= 0.0);
        } else {
            hard = (ic == 1) ? (yn < 0.0) : (yn > 0.0);
        }
        if (hard) a[1] = fmax(a[1], 1.0);
    }
};
void r_cnstr(i64 n, const double *y, const double *ynew, const int *icnstr, int slot)
{
    launch_reduce<0, 2>(RCnstr{y, ynew, icnstr}, n, slot);   // [rdymx, hard-violation flag], both maxima (-DBL_MAX if none)
}

// dcnst0, src/daskr_new.f90:663-713: is the initial y consistent with the constraints?  flag = smallest violating
// index + 1 (INT_MAX if none), as the reference's scan would find it.
__global__ void cnst0_kernel(i64 n, const double *__restrict__ y, const int *__restrict__ icn, int *flag)
{
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int ic = icn[i];
        const double yi = y[i];
        const bool bad = (ic == 2 && yi <= 0.0) || (ic == 1 && yi < 0.0) || (ic == -1 && yi > 0.0) || (ic == -2 && yi >= 0.0);
        if (bad) atomicMin(flag, (int)(i < 2147483646 ? i + 1 : 2147483646));
    }
}
void v_cnst0(i64 n, const double *y, const int *icnstr, int *d_flag)
{
    cnst0_kernel<<<red_grid(n), DKB_RED_BLOCK, 0, g_ctx->stream>>>(n, y, icnstr, d_flag);
    COUNT_LAUNCH();
}

void r_dot(i64 n, const double *x, const double *y, int slot)
{
    launch_reduce<1, 0>(RDot{x, y}, n, slot);
}
void r_dot_s(i64 n, const double *x, double sx, const double *y, int slot)
{
    launch_reduce<1, 0>(RDotS{x, y, sx}, n, slot);
}
void r_dot2(i64 n, const double *vnew, const double *v1, double s1, int slot)
{
    launch_reduce<2, 0>(RDot2{vnew, v1, s1}, n, slot);
}
void r_axpy_dot(i64 n, double *vnew, const double *vi, double si, int hslot, const double *vnext, double sn, int slot)
{
    launch_reduce<1, 0>(RAxpyDot{vnew, vi, g_ctx->d_scal + hslot, vnext, si, sn, 0.0}, n, slot);
}
void r_axpy_nrm(i64 n, double *vnew, const double *vi, double si, int hslot, int slot)
{
    launch_reduce<1, 0>(RAxpyNrm{vnew, vi, g_ctx->d_scal + hslot, si, 0.0}, n, slot);
}
void r_wrms_parts(i64 n, const double *v, const double *rwt, int slot)
{
    launch_reduce<1, 1>(RWrms{v, rwt}, n, slot);
}

// Small device -> host readbacks (scalars, flags) do NOT go through a copy engine: a one-warp kernel stores the words
// into mapped pinned memory and the host waits for the stream.  A cudaMemcpyAsync would queue behind whatever the
// device-to-host DMA engine is doing -- with DKB_OPT_ASYNC_OUT that is the 16 bytes per unknown of the previous
// return's y / yprime, and every scalar of the next step waited for it (measured: 119 ms per step at 4096^2 x 20).
__global__ void mirror_words_kernel(const unsigned *__restrict__ src, unsigned *dst, int n)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}
// hdst: inside a cudaHostAlloc'ed (mapped) buffer; the words are valid when this returns
void fetch_words(const void *dsrc, void *hdst, size_t bytes)
{
    DkbCtx *c = g_ctx;
    void *ddst = nullptr;
    static const bool use_dma = getenv("DKB_FETCH_DMA") != nullptr;   // diagnosis: the old path
    if (use_dma || cudaHostGetDevicePointer(&ddst, hdst, 0) != cudaSuccess) {   // not mapped: the copy engine after all
        cudaGetLastError();
        cudaMemcpyAsync(hdst, dsrc, bytes, cudaMemcpyDeviceToHost, c->stream);
    } else {
        mirror_words_kernel<<<1, 32, 0, c->stream>>>((const unsigned *)dsrc, (unsigned *)ddst, (int)(bytes / 4));
    }
    cudaStreamSynchronize(c->stream);
}

void scal_fetch(int slot, int count, double *out)
{
    DkbCtx *c = g_ctx;
    fetch_words(c->d_scal + slot, c->h_scal + slot, sizeof(double) * count);
    dkb_cuda_check("scal_fetch");
    for (int k = 0; k < count; ++k) out[k] = c->h_scal[slot + k];
}

// ddwnrm, src/daskr_new.f90:828-868.  One pass gives sum (v*rwt)^2 and max|v*rwt|;
// vmax*sqrt(sum((p/vmax)^2)/N) == sqrt(sum(p^2)/N) up to rounding.  The reference's
// max-scaled second pass is only taken when the plain sum over/underflows.
double r_wrms(i64 n, const double *v, const double *rwt)
{
    DkbCtx *c = g_ctx;
    ProfScope ps(PROF_NORM);
    const int slot = 40;
    double r[2];
    r_wrms_parts(n, v, rwt, slot);
    scal_fetch(slot, 2, r);
    double sum = r[0], vmax = r[1];
    if (!(vmax > 0.0)) return (vmax == vmax) ? 0.0 : vmax;   // 0 for the zero vector, NaN propagates
    double nn = (double)c->neq_g;
    if (isfinite(sum) && sum > 1e-280) return sqrt(sum / nn);
    launch_reduce<1, 0>(RWrmsScaled{v, rwt, vmax}, n, slot);
    scal_fetch(slot, 1, r);
    return vmax * sqrt(r[0] / nn);
}

// ddstp's error estimates at orders k, k-1, k-2 (src/daskr_new.f90:360-378) in ONE pass: the weighted RMS norms of
//   e,   phi(:,k+1) + e,   phi(:,k) + (phi(:,k+1) + e)
// -- the reference forms the two sums in `delta` and takes three norms, 96 bytes per unknown; here 32.  nn = 1, 2, 3
// norms (k = 1 needs only the first, k = 2 the first two).  Returns false if a plain sum of squares over/underflowed:
// the caller then takes the reference's statement-by-statement route with the max-scaled norm.
struct RWrms3 {
    const double *e, *rwt, *p1, *p0;
    int nn;
    __device__ void init() {}
    __device__ void operator()(i64 i, double (&a)[6]) const
    {
        const double w = rwt[i], ei = e[i];
        double q = ei * w;
        a[0] = a[0] + q * q;
        a[3] = fmax(a[3], fabs(q));
        if (nn > 1) {
            const double d1 = p1[i] + ei;          // delta = phi(:,kp1) + e (:367)
            q = d1 * w;
            a[1] = a[1] + q * q;
            a[4] = fmax(a[4], fabs(q));
            if (nn > 2) {
                const double d2 = p0[i] + d1;      // delta = phi(:,k) + delta (:374)
                q = d2 * w;
                a[2] = a[2] + q * q;
                a[5] = fmax(a[5], fabs(q));
            }
        }
    }
};
bool r_wrms3(i64 n, int nn, const double *e, const double *rwt, const double *phi_kp1, const double *phi_k, double *out)
{
    DkbCtx *c = g_ctx;
    ProfScope ps(PROF_NORM);
    const int slot = 52;
    launch_reduce<3, 3>(RWrms3{e, rwt, phi_kp1, phi_k, nn}, n, slot);
    double r[6];
    scal_fetch(slot, 6, r);
    const double cnt = (double)c->neq_g;
    for (int k = 0; k < nn; ++k) {
        const double sum = r[k], vmax = r[3 + k];
        if (!(vmax > 0.0)) {
            if (vmax == vmax) out[k] = 0.0;        // zero vector
            else return false;                     // NaN: let the reference path report it
        } else if (isfinite(sum) && sum > 1e-280) {
            out[k] = sqrt(sum / cnt);
        } else {
            return false;
        }
    }
    return true;
}

double r_nrm2(i64 n, const double *x)
{
    double r;
    r_dot(n, x, x, 42);
    scal_fetch(42, 1, &r);
    return sqrt(r);
}
