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
#define GS_U 8    // tile rows whose loads are in flight together

#ifdef GS_DEBUG_CLOCKS
__device__ long long gs_dbg[256 * 8];   // phase clocks of one CTA per launch
#define GS_CLK(k) if (dbg) dbgp[k] = clock64()
#else
#define GS_CLK(k)
#endif

struct GsArgs {
    int mx, my, ty, ntx, nty;   // my = rows of this rank's slab
    int bot_bnd, top_bnd;       // the slab's first / last row is a domain boundary (mirror coefficient gamma2)
    i64 mxns;
    double beta1[DKB_NS_MAX], gamma1[DKB_NS_MAX], dinv[DKB_NS_MAX];   // beta2 = 2*beta1, gamma2 = 2*gamma1 (:452-455)
};

// Spin until a tile's done flag carries this application's epoch (persistent kernel only).
__device__ __forceinline__ void gs_wait(const unsigned *flag, unsigned epoch)
{
    unsigned v;
    for (;;) {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v == epoch) break;
        __nanosleep(64);
    }
}

// One tile of one iteration.  IT1: first iteration (x = dinv*z, no U step, z := 0 + x).
// Shared memory: (ty+2) rows x PITCH doubles; row 0 / point column 0 = lower / left halo of THIS sweep, row hgt+1 /
// point column w+1 = upper / right neighbours of the PREVIOUS sweep (U step).  PITCH = 34*NS + 1: the odd
// pad makes the skewed accesses of phase B (row s-lane, column lane) hit 32 different banks.
template <int NS, int SPC, bool IT1>
__device__ __forceinline__ void gs_tile(const GsArgs &a, int I, int J, int g0, double *__restrict__ z,
                                        double *__restrict__ x, double *sm, int t,
                                        const unsigned *wait_left = nullptr, const unsigned *wait_below = nullptr,
                                        unsigned *edge_flag = nullptr, unsigned epoch = 0, bool last_sweep = false)
{
#ifdef GS_DEBUG_CLOCKS
    const bool dbg = blockIdx.x == gridDim.x / 2 && threadIdx.x == 0 && t < 256;
    long long *const dbgp = gs_dbg + (t & 255) * 8;
#endif
    GS_CLK(0);
    constexpr int PITCH = (GS_TX + 2) * SPC + 1;
    constexpr int UA = GS_U, UC = GS_U;   // g0 = first species of this CTA's group of SPC species
    const int mx = a.mx, my = a.my;
    const i64 mxns = a.mxns;
    const int tid = threadIdx.x;
    const int jx0 = I * GS_TX, jy0 = J * a.ty;
    const int w = min(GS_TX, mx - jx0), hgt = min(a.ty, my - jy0);
    const int rtop = my - 1 - jy0;                      // tile row that is the last mesh row (>= hgt: not in this tile)
    const bool edge = rtop < hgt;

    // Phases A and C: thread tid < w*NS owns offset tid of every tile row (a tile row is w*NS contiguous doubles
    // of the mesh row).  Loads are unconditional and issued GS_U rows at a time.
    const int q = tid / SPC, i = tid - q * SPC;
    const bool mine = tid < w * SPC;
    const int ic = mine ? i : 0;
    const int jx = jx0 + (mine ? q : 0) + 1;            // 1-based mesh column, as in the reference
    const i64 col0 = (i64)jy0 * mxns + (i64)(jx - 1) * NS + g0 + ic;   // this thread's unknown in tile row 0
    double *const smp = sm + PITCH + SPC + tid;                    // ... and its shared-memory slot

    // ---- phase A: right-hand side of the triangular solve for every point of the tile, and the halos ----
    {
        const double b1 = a.beta1[g0 + ic], g1 = a.gamma1[g0 + ic], dinv = a.dinv[g0 + ic];
        const double betU = (jx == 1) ? 2 * b1 : b1;              // (D-inverse)*U*x (:486-530)
        const bool has_right = jx < mx;
        if (IT1) {
            const double *src = z + col0;
            int r0 = 0;
            for (; r0 + UA <= hgt; r0 += UA, src += UA * mxns) {
                double v1[UA];
#pragma unroll
                for (int u = 0; u < UA; ++u) v1[u] = __ldcg(&src[u * mxns]);
#pragma unroll
                for (int u = 0; u < UA; ++u)
                    if (mine) smp[(r0 + u) * PITCH] = dinv * v1[u];       // x = dinv*z (:470-481)
            }
            for (; r0 < hgt; ++r0, src += mxns) {                         // rows of a partial last batch
                const double v1 = __ldcg(&src[0]);
                if (mine) smp[r0 * PITCH] = dinv * v1;
            }
        } else {
            // The previous sweep's x of the tile, of the row above it and of the column to its right go to shared
            // memory once (each value is the right neighbour of one point and the upper neighbour of another), then
            // the U step runs in place, bottom-up in batches of rows: read (registers), barrier, write.
            const bool has_up = jy0 + hgt < my;
            const int nrows = hgt + (has_up ? 1 : 0);
            const double *src = x + col0;
            // the column to the right of the tile: one value per thread (hgt*SPC <= blockDim), in flight with the rest
            const bool has_hr = (jx0 + w < mx) && tid < hgt * SPC;
            double hr = 0.0;
            if (has_hr) hr = __ldcg(&x[(i64)(jy0 + q) * mxns + (i64)(jx0 + w) * NS + g0 + i]);
            constexpr int UL = 17;                                        // 33 rows = two rounds of loads in flight
            for (int r0 = 0; r0 < nrows; r0 += UL) {
                double v1[UL];
#pragma unroll
                for (int u = 0; u < UL; ++u) v1[u] = __ldcg(&src[(i64)min(r0 + u, nrows - 1) * mxns]);
#pragma unroll
                for (int u = 0; u < UL; ++u)
                    if (mine && r0 + u < nrows) smp[(r0 + u) * PITCH] = v1[u];
            }
            if (has_hr) sm[(q + 1) * PITCH + (w + 1) * SPC + i] = hr;
            __syncthreads();
            for (int r0 = 0; r0 < hgt; r0 += UA) {
                double v1[UA], v2[UA];
#pragma unroll
                for (int u = 0; u < UA; ++u) {
                    const int r = min(r0 + u, hgt - 1);
                    v1[u] = smp[r * PITCH + SPC];                          // x(p + 1)
                    v2[u] = smp[(r + 1) * PITCH];                          // x(p + row); unused in the last mesh row
                }
                __syncthreads();
#pragma unroll
                for (int u = 0; u < UA; ++u) {
                    const int r = r0 + u;
                    const double gam = (r == 0 && jy0 == 0 && a.bot_bnd) ? 2 * g1 : g1;
                    const double t1 = betU * v1[u], t2 = gam * v2[u];
                    const bool lasty = edge && r == rtop;
                    const double uu = has_right ? (lasty ? t1 : t1 + t2) : (lasty ? 0.0 : t2);
                    if (mine && r < hgt) smp[r * PITCH] = uu;
                }
            }
        }
    }
    // ---- halos: this sweep's values of the tile to the left and of the tile below.  In the persistent kernel those two
    // tiles are waited for only here: everything above needs the PREVIOUS sweep only, so it is off the critical path.
    if (wait_left || wait_below) {
        if (tid == 0) {
            if (wait_left) gs_wait(wait_left, epoch);
            if (wait_below) gs_wait(wait_below, epoch);
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 2; ++k)                                   // left halo: up to 64 rows = 2 per thread
        if (jx0 > 0 && q + 32 * k < hgt)
            sm[(q + 32 * k + 1) * PITCH + i] = __ldcg(&x[(i64)(jy0 + q + 32 * k) * mxns + (i64)(jx0 - 1) * NS + g0 + i]);
    if (jy0 > 0 && mine) sm[SPC + tid] = __ldcg(&x[col0 - mxns]);

    GS_CLK(1);
    // z of the first UC rows for phase C: in flight while phase B runs
    double zin0[UC];
#pragma unroll
    for (int u = 0; u < UC; ++u) zin0[u] = IT1 ? 0.0 : __ldcg(&z[col0 + min(u, hgt - 1) * mxns]);
    __syncthreads();
    GS_CLK(2);

    // ---- phase B: [(I - (D-inverse)*L)]-inverse (:532-566) ----
    // Warp = species, lane = tile column.  At step s lane q updates tile row s-q: its lower neighbour is its own
    // previous result (a register); its left neighbour is what lane q-1 stored one step earlier (lane 0: the halo
    // column).  The species are independent, so the warps never synchronise with each other.
    {
        const int lane = tid & 31, sp = tid >> 5;                 // blockDim = 32*SPC
        const int jxb = jx0 + lane + 1;
        const double b1 = a.beta1[g0 + sp], g1 = a.gamma1[g0 + sp];
        const double bet = (jxb == mx) ? 2 * b1 : b1;
        const bool has_left = jxb != 1;
        const unsigned nrow = (lane < w) ? (unsigned)hgt : 0u;    // rows this lane works on
        double *p = sm + PITCH + (lane + 1) * SPC + sp;           // row 0 of this lane's column
        double below = p[-PITCH];                                 // lower halo (unused when jy0 == 0)
        const int nsteps = w + hgt - 1;
        if (edge || jy0 == 0) {
            // tiles with mesh row 1 (no lower neighbour) or the last mesh row (gamma2)
            const int r_nob = (jy0 == 0) ? 0 : -1;
            for (int s = 0; s < nsteps; ++s) {
                const int r = s - lane;
                if ((unsigned)r < nrow) {
                    double v = *p;
                    const double left = p[-SPC];
                    const double gam = (r == rtop && a.top_bnd) ? 2 * g1 : g1;
                    if (has_left) v = v + bet * left;
                    if (r != r_nob) v = v + gam * below;
                    *p = v;
                    p += PITCH;
                    below = v;
                }
                __syncwarp();
            }
        } else {
            // (a branch-free body -- idle lanes computing on valid addresses and dropping the result -- was slower:
            // 1.85 vs 1.56 ms at 1024x1024x20)
#pragma unroll 4
            for (int s = 0; s < nsteps; ++s) {
                if ((unsigned)(s - lane) < nrow) {
                    double v = *p;
                    const double left = p[-SPC];
                    if (has_left) v = v + bet * left;
                    v = v + g1 * below;
                    *p = v;
                    p += PITCH;
                    below = v;
                }
                __syncwarp();
            }
        }
    }
    GS_CLK(3);
    __syncthreads();
    GS_CLK(4);
    // Persistent kernel: the right column and the top row are all that the neighbours of THIS sweep need.  Write them
    // first and flag them, so that the tiles to the right and above start their wavefront while phase C streams.
    // (Phase C stores the same doubles to the same addresses once more; a reader that already holds the edge flag sees
    // identical bits either way -- aligned 8-byte stores of equal values.)
    if (edge_flag) {
        for (int e = tid; e < hgt * SPC; e += blockDim.x) {
            const int r = e / SPC, ii = e - r * SPC;
            x[(i64)(jy0 + r) * mxns + (i64)(jx0 + w - 1) * NS + g0 + ii] = sm[(r + 1) * PITCH + w * SPC + ii];
        }
        for (int e = tid; e < w * SPC; e += blockDim.x) {
            const int qq = e / SPC, ii = e - qq * SPC;
            x[(i64)(jy0 + hgt - 1) * mxns + (i64)(jx0 + qq) * NS + g0 + ii] = sm[hgt * PITCH + (qq + 1) * SPC + ii];
        }
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(edge_flag), "r"(epoch) : "memory");
        }
    }

    // ---- phase C: x out, z := z + x (:568-570; z was zeroed before the first iteration, :478) ----
    {
        double *xo = x + col0, *zo = z + col0;
        const double *sp = smp;
        int r0 = 0;
        for (; r0 + UC <= hgt; r0 += UC, xo += UC * mxns, zo += UC * mxns, sp += UC * PITCH) {
            double zin[UC];
            const bool more = !IT1 && r0 + UC < hgt;            // prefetch the next batch of z before storing this one
#pragma unroll
            for (int u = 0; u < UC; ++u) zin[u] = more ? __ldcg(&zo[min(UC + u, hgt - 1 - r0) * mxns]) : 0.0;
            if (mine) {
#pragma unroll
                for (int u = 0; u < UC; ++u) {
                    const double v = sp[u * PITCH];
                    if (!last_sweep) xo[u * mxns] = v;   // nobody reads x after the last sweep (the edges are out already)
                    zo[u * mxns] = zin0[u] + v;
                }
            }
#pragma unroll
            for (int u = 0; u < UC; ++u) zin0[u] = zin[u];
        }
        if (mine) {
#pragma unroll
            for (int u = 0; u < UC; ++u) {
                if (r0 + u < hgt) {
                    const double v = sp[u * PITCH];
                    if (!last_sweep) xo[u * mxns] = v;   // nobody reads x after the last sweep (the edges are out already)
                    zo[u * mxns] = zin0[u] + v;
                }
            }
        }
    }
    GS_CLK(5);
}

template <int NS, int SPC>
__global__ void __launch_bounds__(32 * SPC) gs_step_kernel(const __grid_constant__ GsArgs a, int t, double *__restrict__ z,
                                                           double *__restrict__ x)
{
    extern __shared__ double sm[];
    // ---- which (iteration, tile, species group) is this CTA? ----
    constexpr int NG = NS / SPC;
    const int g0 = (blockIdx.x % NG) * SPC;
    int b = blockIdx.x / NG, iter = 0, I = 0, J = 0;
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
    if (iter == 1) gs_tile<NS, SPC, true>(a, I, J, g0, z, x, sm, t);
    else gs_tile<NS, SPC, false>(a, I, J, g0, z, x, sm, t);
}

// ---- persistent variant: one launch per application ----
// The tiles are numbered in the order of the launch schedule above (time step, then sweep, then column), and the
// CTAs of ONE launch take them from an atomic ticket counter.  A tile (I,J) of sweep k needs
//   (I+1,J), (I,J+1) and (I,J) of sweep k-1, complete ("done" flag)      -- before its U step (phase A),
//   the right column of (I-1,J) and the top row of (I,J-1) of sweep k    -- before its wavefront (phase B);
// the second pair is waited for only after phase A, and a tile publishes its right column and top row ("edge" flag)
// right after its wavefront, before phase C streams the rest: along the chain of dependent tiles only the halo
// load, the wavefront and the edge store are serial.  All dependencies have smaller numbers, so they have been taken
// by CTAs that are running (a CTA only takes a ticket when it is resident): the tile with the smallest unfinished
// number never waits, and the sweep cannot deadlock whatever the number of resident CTAs.  Write-after-read: a
// tile overwrites x in phase C, after it has seen the edge flags of (I-1,J) and (I,J-1) -- the only readers of its
// old values, in their phase A, which precedes their edge flag.
// Loads of x and z go to L2 (__ldcg): L1 is not coherent between SMs inside a launch.
struct GsSched {
    unsigned *ticket;          // next tile number
    unsigned *done;            // [GS_ITMAX][nty][ntx]: == epoch when the tile of this application is finished
    unsigned *edge;            // same shape: == epoch when the tile's right column and top row are in memory
    const int *first;          // first[t] = number of tiles before time step t, first[nt] = total
    unsigned epoch;
    int nt, total;
};

template <int NS>
__global__ void __launch_bounds__(32 * NS) gs_persistent_kernel(const __grid_constant__ GsArgs a, const GsSched s,
                                                               double *__restrict__ z, double *__restrict__ x)
{
    extern __shared__ double sm[];
    __shared__ int cur[4];     // iter, I, J, time step (iter = 0: no tiles left)
    int t = 0;                 // thread 0's cursor into first[]: ticket numbers only grow
    for (;;) {
        if (threadIdx.x == 0) {
            const unsigned n = atomicAdd(s.ticket, 1u);
            int iter = 0, I = 0, J = 0;
            if (n < (unsigned)s.total) {
                while ((int)n >= s.first[t + 1]) ++t;
                int b = (int)n - s.first[t];
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
                const unsigned *own = s.done + (size_t)(iter - 1) * a.nty * a.ntx;
                const unsigned *prv = own - (size_t)a.nty * a.ntx;
                // (I-1,J) and (I,J-1) of this sweep are waited for inside the tile, after its U step
                if (iter > 1 && I < a.ntx - 1) gs_wait(prv + (size_t)J * a.ntx + I + 1, s.epoch);
                if (iter > 1 && J < a.nty - 1) gs_wait(prv + (size_t)(J + 1) * a.ntx + I, s.epoch);
                if (iter > 1) gs_wait(prv + (size_t)J * a.ntx + I, s.epoch);   // the tile's own previous sweep (x, z in place)
            }
            cur[0] = iter; cur[1] = I; cur[2] = J; cur[3] = t;
        }
        __syncthreads();
        const int iter = cur[0], I = cur[1], J = cur[2], tt = cur[3];
        if (iter == 0) return;
        unsigned *own = s.edge + (size_t)(iter - 1) * a.nty * a.ntx;
        const unsigned *wl = I > 0 ? own + (size_t)J * a.ntx + I - 1 : nullptr;
        const unsigned *wb = J > 0 ? own + (size_t)(J - 1) * a.ntx + I : nullptr;
        unsigned *ef = own + (size_t)J * a.ntx + I;
        if (iter == 1) gs_tile<NS, NS, true>(a, I, J, 0, z, x, sm, tt, wl, wb, ef, s.epoch);
        else gs_tile<NS, NS, false>(a, I, J, 0, z, x, sm, tt, wl, wb, ef, s.epoch, iter == GS_ITMAX);
        __syncthreads();       // every thread's stores are issued; shared memory is free for the next tile
        if (threadIdx.x == 0) {
            __threadfence();
            unsigned *flag = s.done + ((size_t)(iter - 1) * a.nty + J) * a.ntx + I;
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(s.epoch) : "memory");
        }
    }
}

template <int NS, int SPC>
static void gs_launch(const GsArgs &a, int ntile, int t, double *z, double *x, cudaStream_t st)
{
    const size_t smem = ((size_t)(GS_TX + 2) * SPC + 1) * sizeof(double) * (a.ty + 2);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        cudaFuncSetAttribute(gs_step_kernel<NS, SPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        smem_set = smem;
    }
    gs_step_kernel<NS, SPC><<<ntile * (NS / SPC), 32 * SPC, smem, st>>>(a, t, z, x);
}

// Tuning notes (ns = 20, B200; profiles/README.md has the numbers).  One 32x32 tile costs ~20 us on its SM (30 us when
// all SMs stream at once): 4 dependent rounds of loads (phase A), 63 wavefront steps of ~170 clk (phase B: LDS -> DMUL
// -> DADD -> DADD -> STS; DADD/DMUL 9 clk, STS->LDS 34 clk, the rest is issue: 20 warps x 17 instructions per step) and
// 4 rounds of loads/stores (phase C); issue slots are 1/3 used, the dominant stall is long-scoreboard.
//   one launch per time step : 1.74 ms (1024^2), 19.9 ms (4096^2)   -- first version, DKB_GS_PER_STEP=1
//   persistent scheduler     : 1.56 ms, 16.9 ms
//   + late waits, early edges: 1.2 ms (2.8 TB/s algorithmic), 17.4 ms (3.1 TB/s); per tile A ~15k, B ~11k, C ~12k clk
//   + x of the previous sweep staged once in shared memory, U step in place: 1.1 ms, 14.7 ms (3.7 TB/s)
// Tried without gain: two CTAs per tile with half of the species each (same number of warps per SM), 16- and 32-row
// load batches (register pressure at 640 threads), a branch-free wavefront body (idle lanes computing costs more than
// the branch).  The next step is TMA double-buffering inside the persistent CTA so that the loads of the next tile
// and the stores of the previous one overlap the wavefront.
// z := (approximately) A_S^-1 z, x = N words of scratch.
// Several ranks: the sweep is lexicographic, so slab r would have to wait for slab r-1.  Instead every rank sweeps
// its own slab extended by GS_OV_BELOW rows of the slab below and GS_OV_ABOVE rows of the slab above (one exchange
// of those rows of z per application) and keeps its own rows.  Why that reproduces the global sweep: the triangular
// solve couples a point to the rows below it through paths of total weight <= (gamma/(1-beta))^m over m rows, and
// beta + gamma < 1/2 -- (1/3)^m on a square mesh, at worst (1/2)^m on a very anisotropic one; the U step reaches
// one row upwards per iteration, four rows in all.  Own rows agree with the single-GPU sweep to ~1e-13 relative on
// a square mesh (tests/test_oracle_cpu.py checks this on the matrix), not bit for bit.
#define GS_OV_BELOW DKB_GS_OV_BELOW
#define GS_OV_ABOVE DKB_GS_OV_ABOVE
static int gs_sweep(double hl0, int my, bool bot_bnd, bool top_bnd, double *z, double *x, bool may_cut = false, bool *moved = nullptr);

// z := (approximately) A_S^-1 z.  *res = where the result is: z itself, or x when the sweep ran out of place (several
// ranks: column panels, gs_stream.cu).  Callers that can read the result from either array (datv, dspigm) use this
// form; web_gauss_seidel() below keeps the in-place contract of the psol callback.
int web_gauss_seidel_to(double hl0, double *z, double *x, double **res)
{
    DkbCtx *c = g_ctx;
    const WebPar &wp = c->web;
    *res = z;
    if (c->nranks <= 1) {
        // DKB_GS_FORCE_PANELS=1 (tests): the column-panel decomposition of the multi-rank path on one GPU
        const char *fp = getenv("DKB_GS_FORCE_PANELS");
        if (fp && atoi(fp) == 1) {
            bool moved = false;
            const int rc = gs_sweep(hl0, c->my, true, true, z, x, true, &moved);
            if (moved) *res = x;
            return rc;
        }
        return gs_sweep(hl0, c->my, true, true, z, x);
    }
    const i64 mxns = (i64)wp.mx * wp.ns;
    const int my_min = wp.my / c->nranks;                    // every slab has at least this many rows
    const int ob = std::min(GS_OV_BELOW, my_min), oa = std::min(GS_OV_ABOVE, my_min);
    const int nb = c->rank > 0 ? ob : 0, na = c->rank < c->nranks - 1 ? oa : 0;
    const i64 n_own = (i64)c->my * mxns;
    // The solver's vectors carry DKB_GS_OV_BELOW rows of room on both sides (alloc_solver): sweep there, the overlap
    // rows arrive straight in that room.  Any other caller's arrays go through a copy.
    auto in_arena = [&](const double *p) { return c->arena && p >= c->arena && p < c->arena + c->arena_doubles; };
    if (in_arena(z) && in_arena(x) && c->halo_pad >= (i64)ob * mxns) {
        slab_overlap_exchange(z, n_own, (i64)ob * mxns, (i64)oa * mxns, z - (i64)nb * mxns, z + n_own);
        bool moved = false;
        const int rc = gs_sweep(hl0, nb + c->my + na, c->rank == 0, c->rank == c->nranks - 1, z - (i64)nb * mxns, x - (i64)nb * mxns,
                                true, &moved);
        if (moved) *res = x;
        return rc;
    }
    const i64 need = (i64)(nb + c->my + na) * mxns;
    if (need > c->gs_cap) {
        cudaFree(c->gs_z);
        cudaFree(c->gs_x);
        c->gs_z = c->gs_x = nullptr;
        c->gs_cap = 0;
        const i64 cap = (i64)(c->my + 1 + ob + oa) * mxns;
        if (cudaMalloc(&c->gs_z, sizeof(double) * cap) != cudaSuccess || cudaMalloc(&c->gs_x, sizeof(double) * cap) != cudaSuccess)
            return dkb_fail(DKB_E_CUDA, "Gauss-Seidel factor: cannot allocate the overlapped-slab work vectors");
        c->gs_cap = cap;
    }
    double *zo = c->gs_z + (i64)nb * mxns;
    cudaMemcpyAsync(zo, z, sizeof(double) * n_own, cudaMemcpyDeviceToDevice, c->stream);
    slab_overlap_exchange(z, n_own, (i64)ob * mxns, (i64)oa * mxns, c->gs_z, zo + n_own);
    bool moved = false;
    const int rc = gs_sweep(hl0, nb + c->my + na, c->rank == 0, c->rank == c->nranks - 1, c->gs_z, c->gs_x, true, &moved);
    cudaMemcpyAsync(z, (moved ? c->gs_x : c->gs_z) + (i64)nb * mxns, sizeof(double) * n_own, cudaMemcpyDeviceToDevice, c->stream);
    return rc;
}

int web_gauss_seidel(double hl0, double *z, double *x)
{
    DkbCtx *c = g_ctx;
    double *res = z;
    const int rc = web_gauss_seidel_to(hl0, z, x, &res);
    if (rc == 0 && res != z)
        cudaMemcpyAsync(z, res, sizeof(double) * (size_t)c->my * c->web.mx * c->web.ns, cudaMemcpyDeviceToDevice, c->stream);
    return rc;
}

// The sweep over `my` rows of z (in place), x = scratch of the same size.  bot_bnd / top_bnd: the first / last row is a
// domain boundary (mirror coefficients beta2/gamma2 as in the reference); otherwise the neighbour row is simply absent.
// may_cut (several ranks: the slab sweep is an overlapped decomposition already): the streaming kernel may cut the columns
// into overlapped panels as well, which needs separate in / out arrays: the result is then in x and *moved = true.
static int gs_sweep(double hl0, int my, bool bot_bnd, bool top_bnd, double *z, double *x, bool may_cut, bool *moved)
{
    DkbCtx *c = g_ctx;
    ProfScope ps(PROF_GS);
    // Two kernels.  gs_stream.cu makes ONE pass over the slab with all five iterations fused (16 bytes per unknown); its
    // warps form a pipeline across the mx/32 column strips, which takes ~50 rows per strip to fill.  The five-pass tile
    // scheduler below moves 160 bytes per unknown but has no such start-up.  Measured on B200 at mx = 4096, ns = 20
    // (profiles/README.md): 544 rows 2.9 vs 2.1 ms, 1056 rows 3.3 vs 3.6 ms, 4096 rows 6.7 vs 14.3 ms -- so short, wide
    // slabs (8 GPUs on the 4096^2 mesh) keep the tile scheduler.  DKB_GS_STREAM = 0 / 1 forces one of them (tests, A/B runs).
    {
        const char *st_env = getenv("DKB_GS_STREAM");
        const int ns0 = c->web.ns;
        const bool tiles_ok = ns0 == 2 || ns0 == 4 || ns0 == 6 || ns0 == 8 || ns0 == 10 || ns0 == 12 || ns0 == 16 || ns0 == 20 || ns0 == 24 || ns0 == 32;
        // column panels of ~512 columns (several ranks only; DKB_GS_PANEL_COLS overrides, 0 = none)
        const char *pc_env = getenv("DKB_GS_PANEL_COLS");
        const int pcols = pc_env ? atoi(pc_env) : 512;
        const int npanels = (may_cut && pcols > 0) ? std::max(1, (c->web.mx + pcols / 2) / pcols) : 1;
        bool stream = !tiles_ok || npanels > 1 || 5 * (i64)my >= (i64)c->web.mx;
        if (st_env) stream = atoi(st_env) != 0 || !tiles_ok;
        if (stream) {
            if (npanels > 1) {
                if (moved) *moved = true;
                return gs_stream_sweep(hl0, my, bot_bnd, top_bnd, z, x, npanels);
            }
            return gs_stream_sweep(hl0, my, bot_bnd, top_bnd, z, z, 1);
        }
    }
    const WebPar &wp = c->web;
    GsArgs a;
    a.mx = wp.mx;
    a.my = my;
    a.bot_bnd = bot_bnd;
    a.top_bnd = top_bnd;
    a.mxns = (i64)wp.mx * wp.ns;
    for (int i = 0; i < wp.ns; ++i) {   // :447-457
        const double elamda = 1.0 / (1.0 + 2 * hl0 * (wp.cox[i] + wp.coy[i]));
        a.beta1[i] = hl0 * wp.cox[i] * elamda;
        a.gamma1[i] = hl0 * wp.coy[i] * elamda;
        a.dinv[i] = elamda;
    }
    const int spc = wp.ns;   // species per CTA (a split over several CTAs was tried, see the notes above)
    // tile height: at most 32 rows, within 200 KB of the 227 KB of shared memory
    const size_t rowbytes = ((size_t)(GS_TX + 2) * spc + 1) * sizeof(double);
    // Tile height: 32 rows (fewest halo bytes, shortest chain of tiles).  DKB_GS_TY overrides (tests, tuning).
    const char *ty_env = getenv("DKB_GS_TY");
    const int ty_want = ty_env ? std::max(1, std::min(32, atoi(ty_env))) : 32;
    int ty = (int)std::min<size_t>(ty_want, (200u * 1024u) / rowbytes - 2);
    ty = std::max(1, std::min(ty, a.my));
    a.ty = ty;
    a.ntx = (wp.mx + GS_TX - 1) / GS_TX;
    a.nty = (a.my + ty - 1) / ty;
    const int ndiag = a.ntx + a.nty - 1;
    const int nt = ndiag + 2 * (GS_ITMAX - 1);
    auto tiles_at = [&](int t) {
        int ntile = 0;
        for (int k = 0; k < GS_ITMAX; ++k) {
            const int d = t - 2 * k;
            if (d < 0 || d > ndiag - 1) continue;
            ntile += std::min(a.ntx - 1, d) - std::max(0, d - (a.nty - 1)) + 1;
        }
        return ntile;
    };
    // DKB_GS_PER_STEP=1: one launch per time step (the first version; kept for A/B runs and as a cross-check)
    const char *ps_env = getenv("DKB_GS_PER_STEP");
    const int per_step = (ps_env && atoi(ps_env) == 1) ? 1 : 0;
#define GS_CASES(M) \
    switch (wp.ns) { \
        M(2) M(4) M(6) M(8) M(10) M(12) M(16) M(20) M(24) M(32) \
        default: return dkb_fail(DKB_E_UNSUPPORTED, "Gauss-Seidel factor: ns=%d not instantiated (2,4,6,8,10,12,16,20,24,32)", wp.ns); \
    }
    if (!per_step) {
        // ---- one persistent launch: scheduler state (re)built when the tiling changes ----
        if (c->gs_ntx != a.ntx || c->gs_nty != a.nty || !c->gs_done) {
            cudaFree(c->gs_ticket);
            cudaFree(c->gs_done);
            cudaFree(c->gs_first);
            c->gs_ticket = c->gs_done = nullptr;
            c->gs_first = nullptr;
            std::vector<int> first(nt + 1, 0);
            for (int t = 0; t < nt; ++t) first[t + 1] = first[t] + tiles_at(t);
            const size_t nflag = (size_t)GS_ITMAX * a.ntx * a.nty;
            if (cudaMalloc(&c->gs_ticket, sizeof(unsigned)) != cudaSuccess || cudaMalloc(&c->gs_done, sizeof(unsigned) * 2 * nflag) != cudaSuccess ||
                cudaMalloc(&c->gs_first, sizeof(int) * (nt + 1)) != cudaSuccess)
                return dkb_fail(DKB_E_CUDA, "Gauss-Seidel factor: cannot allocate the scheduler state");
            cudaMemsetAsync(c->gs_done, 0, sizeof(unsigned) * 2 * nflag, c->stream);   // done flags, then edge flags
            cudaMemcpyAsync(c->gs_first, first.data(), sizeof(int) * (nt + 1), cudaMemcpyHostToDevice, c->stream);
            cudaStreamSynchronize(c->stream);   // `first` is a local
            c->gs_ntx = a.ntx;
            c->gs_nty = a.nty;
            c->gs_nt = nt;
            c->gs_total = first[nt];
            c->gs_epoch = 0;
        }
        c->gs_epoch += 1;
        if (c->gs_epoch == 0) {   // wrapped: flags of 2^32 applications ago could alias
            cudaMemsetAsync(c->gs_done, 0, sizeof(unsigned) * 2 * (size_t)GS_ITMAX * a.ntx * a.nty, c->stream);
            c->gs_epoch = 1;
        }
        cudaMemsetAsync(c->gs_ticket, 0, sizeof(unsigned), c->stream);
        GsSched sch{c->gs_ticket, c->gs_done, c->gs_done + (size_t)GS_ITMAX * a.ntx * a.nty, c->gs_first, c->gs_epoch, c->gs_nt, c->gs_total};
        const size_t smem = rowbytes * (a.ty + 2);
        static int nsm = 0;
        if (!nsm) cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device);
        // as many CTAs as are resident at once (shared memory decides), fewer if there are fewer tiles
#define GS_P(n) case n: { \
            static size_t set = 0; \
            static int per_sm = 1; \
            if (smem != set) { \
                cudaFuncSetAttribute(gs_persistent_kernel<n>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
                cudaFuncSetAttribute(gs_persistent_kernel<n>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gs_persistent_kernel<n>, 32 * n, smem); \
                per_sm = std::max(1, per_sm); \
                set = smem; \
            } \
            gs_persistent_kernel<n><<<std::min(c->gs_total, nsm * per_sm), 32 * n, smem, c->stream>>>(a, sch, z, x); } break;
        GS_CASES(GS_P)
#undef GS_P
        COUNT_LAUNCH();
    } else {
        for (int t = 0; t < nt; ++t) {
            const int ntile = tiles_at(t);
            if (ntile == 0) continue;
#define GS_L(n) case n: gs_launch<n, n>(a, ntile, t, z, x, c->stream); break;
            GS_CASES(GS_L)
#undef GS_L
            COUNT_LAUNCH();
        }
    }
#undef GS_CASES
    dkb_cuda_check("web_gauss_seidel");
    return 0;
}

#ifdef GS_DEBUG_CLOCKS
extern "C" void dkb_debug_gs_clocks(long long *out)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, gs_dbg, sizeof(long long) * 256 * 8);
}
#endif
