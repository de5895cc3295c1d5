// gs.cu -- the food-web example's spatial preconditioner factor on the device
// (gauss_seidel, example/example_web.f90:393-575).
//
// The reference runs itmax = 5 lexicographic Gauss-Seidel iterations on A_S = I - hl0*J_d, species
// by species decoupled.  Each iteration is
//   (U)  x(p) := beta*x(p+1) + gamma*x(p+row)                      (iterations 2..5; reads values AHEAD
//                                                                   in sweep order, i.e. the previous iterate)
//   (L)  x(p) := (x(p) + beta*x(p-1)) + gamma*x(p-row)             (sweep order: needs the NEW left / lower values)
//   (S)  z    := z + x
// (L) is a two-dimensional recurrence.  Here the mesh is cut into TX x TY tiles; tile (I,J) of
// iteration k only needs tiles (I-1,J), (I,J-1) of the same iteration and tiles (I+1,J), (I,J+1) of
// iteration k-1, so iteration k can work on tile diagonal d = I+J at time step t = d + 2(k-1).  One
// launch per time step processes up to five diagonals (one per iteration) -- diagonals of equal parity
// are never adjacent, so the in-place update of x is race free -- and inside a tile the recurrence
// runs as a wavefront over anti-diagonals of mesh points in shared memory.  Every point is updated
// by the reference's expression on the reference's operands: the result is bit-identical to the
// sequential sweep.  No spin-waits: the dependency order is the launch order on one stream.
#include "dkb_internal.h"

#define GS_ITMAX 5
#define GS_TX 32

struct GsArgs {
    int ns, mx, my, ty, ntx, nty;
    i64 mxns;
    double beta1[DKB_NS_MAX], gamma1[DKB_NS_MAX], dinv[DKB_NS_MAX];   // beta2 = 2*beta1, gamma2 = 2*gamma1 (:452-455)
};

__global__ void __launch_bounds__(1024) gs_step_kernel(const __grid_constant__ GsArgs a, int t, double *__restrict__ z,
                                                       double *__restrict__ x)
{
    extern __shared__ double sm[];   // (ty+1) rows x (GS_TX+1) columns x ns; row 0 / column 0 = lower / left halo
    const int ns = a.ns, mx = a.mx, my = a.my;
    const i64 mxns = a.mxns;

    // ---- which (iteration, tile) is this CTA? ----
    int b = blockIdx.x, iter = 0, I = 0, J = 0;
    for (int k = 0; k < GS_ITMAX; ++k) {
        const int d = t - 2 * k;
        if (d < 0 || d > a.ntx + a.nty - 2) continue;
        const int ilo = max(0, d - (a.nty - 1)), ihi = min(a.ntx - 1, d);
        const int cnt = ihi - ilo + 1;
        if (b < cnt) {
            iter = k + 1;
            I = ilo + b;
            J = d - I;
            break;
        }
        b -= cnt;
    }
    if (iter == 0) return;
    const int jx0 = I * GS_TX, jy0 = J * a.ty;
    const int w = min(GS_TX, mx - jx0), hgt = min(a.ty, my - jy0);
    const int pitch = (GS_TX + 1) * ns;
    const int wns = w * ns;

    // ---- phase A: right-hand side of the triangular solve for every point of the tile, and the halos ----
    for (int e = threadIdx.x; e < hgt * wns; e += blockDim.x) {
        const int r = e / wns, rem = e - r * wns;
        const int q = rem / ns, i = rem - q * ns;
        const int jx = jx0 + q + 1, jy = jy0 + r + 1;   // 1-based mesh indices, as in the reference
        const i64 idx = (i64)(jy - 1) * mxns + (i64)(jx - 1) * ns + i;
        double u;
        if (iter == 1) {
            u = a.dinv[i] * z[idx];                      // x = dinv*z (:470-481)
        } else {                                         // (D-inverse)*U*x (:486-530)
            const double gam = (jy == 1) ? 2 * a.gamma1[i] : a.gamma1[i];
            const bool lasty = (jy == my);
            if (jx == 1) {
                const double b2 = 2 * a.beta1[i];
                u = lasty ? b2 * x[idx + ns] : b2 * x[idx + ns] + gam * x[idx + mxns];
            } else if (jx < mx) {
                u = lasty ? a.beta1[i] * x[idx + ns] : a.beta1[i] * x[idx + ns] + gam * x[idx + mxns];
            } else {
                u = lasty ? 0.0 : gam * x[idx + mxns];
            }
        }
        sm[(r + 1) * pitch + (q + 1) * ns + i] = u;
    }
    if (jx0 > 0)   // left halo: this iteration's values of the tile to the left
        for (int e = threadIdx.x; e < hgt * ns; e += blockDim.x) {
            const int r = e / ns, i = e - r * ns;
            sm[(r + 1) * pitch + i] = x[(i64)(jy0 + r) * mxns + (i64)(jx0 - 1) * ns + i];
        }
    if (jy0 > 0)   // lower halo
        for (int e = threadIdx.x; e < wns; e += blockDim.x) sm[ns + e] = x[(i64)(jy0 - 1) * mxns + (i64)jx0 * ns + e];
    __syncthreads();

    // ---- phase B: [(I - (D-inverse)*L)]-inverse (:532-566), wavefront over the tile's anti-diagonals ----
    {
        const int q = threadIdx.x / ns, i = threadIdx.x - q * ns;
        const bool mine = threadIdx.x < wns;
        const int jx = jx0 + q + 1;
        const double bet = (jx == mx) ? 2 * a.beta1[mine ? i : 0] : a.beta1[mine ? i : 0];
        const double g1 = a.gamma1[mine ? i : 0];
        for (int s = 0; s < w + hgt - 1; ++s) {
            const int r = s - q;
            if (mine && r >= 0 && r < hgt) {
                const int jy = jy0 + r + 1;
                double *own = &sm[(r + 1) * pitch + (q + 1) * ns + i];
                if (!(jy == 1 && jx == 1)) {
                    const double gam = (jy == my) ? 2 * g1 : g1;
                    double v = *own;
                    if (jy == 1) v = v + bet * own[-ns];
                    else if (jx == 1) v = v + gam * own[-pitch];
                    else v = (v + bet * own[-ns]) + gam * own[-pitch];
                    *own = v;
                }
            }
            __syncthreads();
        }
    }

    // ---- phase C: x out, z := z + x (:568-570; z was zeroed before the first iteration, :478) ----
    for (int e = threadIdx.x; e < hgt * wns; e += blockDim.x) {
        const int r = e / wns, rem = e - r * wns;
        const i64 idx = (i64)(jy0 + r) * mxns + (i64)jx0 * ns + rem;
        const double v = sm[(r + 1) * pitch + ns + rem];
        x[idx] = v;
        const double zin = (iter == 1) ? 0.0 : z[idx];
        z[idx] = zin + v;
    }
}

// z := (approximately) A_S^-1 z, x = N words of scratch.  Single rank only: the sweep order couples the slabs.
int web_gauss_seidel(double hl0, double *z, double *x)
{
    DkbCtx *c = g_ctx;
    const WebPar &wp = c->web;
    if (c->nranks > 1) return dkb_fail(DKB_E_ARG, "Gauss-Seidel factor (jpre = 2, 3, 4) is single-GPU in this version");
    if (wp.ns > DKB_NS_MAX || GS_TX * wp.ns > 1024)
        return dkb_fail(DKB_E_ARG, "Gauss-Seidel factor: ns = %d too large (ns <= %d)", wp.ns, 1024 / GS_TX);
    GsArgs a;
    a.ns = wp.ns;
    a.mx = wp.mx;
    a.my = wp.my;
    a.mxns = (i64)wp.mx * wp.ns;
    for (int i = 0; i < wp.ns; ++i) {   // :447-457
        const double elamda = 1.0 / (1.0 + 2 * hl0 * (wp.cox[i] + wp.coy[i]));
        a.beta1[i] = hl0 * wp.cox[i] * elamda;
        a.gamma1[i] = hl0 * wp.coy[i] * elamda;
        a.dinv[i] = elamda;
    }
    // tile height from the shared-memory budget (200 KB of the 227 KB per CTA)
    const size_t colbytes = (size_t)(GS_TX + 1) * wp.ns * sizeof(double);
    int ty = (int)std::min<size_t>(32, (200u * 1024u) / colbytes - 1);
    ty = std::max(1, std::min(ty, wp.my));
    a.ty = ty;
    a.ntx = (wp.mx + GS_TX - 1) / GS_TX;
    a.nty = (wp.my + ty - 1) / ty;
    const size_t smem = colbytes * (ty + 1);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaFuncSetAttribute(gs_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        smem_set = smem;
    }
    const int threads = std::min(1024, ((std::min(GS_TX, wp.mx) * wp.ns + 31) / 32) * 32);
    const int ndiag = a.ntx + a.nty - 1;
    for (int t = 0; t < ndiag + 2 * (GS_ITMAX - 1); ++t) {
        int nblk = 0;
        for (int k = 0; k < GS_ITMAX; ++k) {
            const int d = t - 2 * k;
            if (d < 0 || d > ndiag - 1) continue;
            nblk += std::min(a.ntx - 1, d) - std::max(0, d - (a.nty - 1)) + 1;
        }
        if (nblk == 0) continue;
        gs_step_kernel<<<nblk, threads, smem, c->stream>>>(a, t, z, x);
        COUNT_LAUNCH();
    }
    dkb_cuda_check("web_gauss_seidel");
    return 0;
}
