// gs_stream.cu -- the food-web example's spatial preconditioner factor (gauss_seidel,
// example/example_web.f90:393-575) with ALL FIVE Gauss-Seidel iterations fused into one pass over the mesh.
//
// The reference runs itmax = 5 lexicographic iterations on A_S = I - hl0*J_d, species by species:
//   (U)  x(p) := beta*x(p+1) + gamma*x(p+row)                     iterations 2..5; previous iterate (ahead in sweep order)
//   (L)  x(p) := (x(p) + beta*x(p-1)) + gamma*x(p-row)            needs the NEW left / lower values
//   (S)  z    := z + x
// Iteration k at point (c,r) needs iteration k at (c-1,r), (c,r-1) and iteration k-1 at (c+1,r), (c,r+1): nothing
// else.  gs.cu runs the five iterations as five passes (x and z read and written five times: 160 bytes per unknown).
// Here each value is produced once and consumed from registers:
//
//   * a WARP owns one species on a strip of 32 mesh columns and streams up the rows; lane l works on column
//     32*I - (k-1) + l in iteration k (the strip shifts one column to the left per iteration, so the right-hand
//     neighbour of iteration k-1 is always in the same lane);
//   * at step s lane l updates row s - l - (k-1) of iteration k, for all five k at once.  With that skew every
//     operand is either the lane's own result of the previous step (lower neighbour; right neighbour of the previous
//     iterate) or the result of lane l-1 of the previous step (left neighbour; upper neighbour of the previous
//     iterate): one __shfl_up per iteration and step, five independent dependency chains per warp;
//   * the running sum z = ((((0 + x1) + x2) + x3) + x4) + x5 follows the mesh point from lane to lane the same way
//     (the point visited by lane l in iteration k is visited by lane l+1 in iteration k+1, two steps later);
//   * lane 0's left-hand operands come from the strip to the left, which runs >= 48 steps ahead and leaves its lane-31
//     results (5 iterates + 4 partial sums per row) in a global exchange buffer (L2 resident); a persistent CTA takes
//     (strip, species group) items in strip order from a ticket counter, so a CTA only ever waits for a CTA that is
//     already running (no deadlock for any number of resident CTAs);
//   * z enters and leaves through a 64-row circular window in shared memory, loaded and stored in coalesced batches of
//     8 rows by the whole CTA (SPC species of 32 points = whole 32-byte sectors), in place: a point's input slot is
//     reused for its result.
// Traffic: z read once and written once (16 bytes per unknown) + 2*10/32 words per unknown of strip exchange, instead
// of 160.  Every point is updated by the reference's expression on the reference's operands in the reference's order,
// so the result is bit-identical to the sequential sweeps (tests/test_gpu_gs.py).
#include "dkb_internal.h"

#ifndef GSS_B
#define GSS_B 8        // steps (= rows entering the window) per batch
#endif
#define GSS_WIN 64     // rows of the circular z window (rows s0-44 .. s0+15 are live during the batch that starts at step s0;
                       // 60 would do and lets three CTAs fit an SM, but three are slower than two -- see gss_launch)
#define GSS_ROW(r) ((r) & (GSS_WIN - 1))
#define GSS_HS 10      // doubles per exchange slot: x1..x5, sum1..sum4 of ONE step of the producer's lane 31, pad (80 bytes, 16-byte aligned)
#ifdef GSS_DEBUG
__device__ long long gss_dbg[4096 * 8];   // phase clocks of one CTA (the item in the middle of the schedule), per batch
#define GSS_CLK(k) if (dbg && tid == 0) gss_dbg[dbgb * 8 + (k)] = clock64()
#else
#define GSS_CLK(k)
#endif
#ifdef GSS_WEAK_LD
#define GSS_LDZ(p) (*(p))          // z is read once per application and was written by an earlier kernel: any cache will do
#else
#define GSS_LDZ(p) __ldcg(p)
#endif
#define GSS_EMPTY (-1LL)   // bit pattern of an exchange word that has not been written: all ones (cudaMemset 0xFF), a NaN payload
                           // that no floating-point instruction returns (results carry the canonical NaN)

struct GssArgs {
    int mx, H, nstrips, ng;      // H = rows swept; nstrips = strips per panel
    int npanels, pw;             // column panels (see gs_stream_sweep): panel p owns columns [p*pw, (p+1)*pw)
    int ovl, ovr;                // columns of the neighbouring panels swept redundantly on the left / right
    int bot_bnd, top_bnd;        // first / last row is a domain boundary (mirror coefficient gamma2 = 2*gamma1)
    int s0_last;                 // first step of the last batch (the window is flushed by then)
    int edge_fast;               // 0: the edge strips take the fully general step (tuning switch)
    unsigned poll_ns;            // sleep between two polls of the exchange words
    i64 hp;                      // exchange slots per (strip boundary, species)
    double beta1[DKB_NS_MAX], gamma1[DKB_NS_MAX], dinv[DKB_NS_MAX];
};
__device__ __forceinline__ double gss_poll(const double *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return __longlong_as_double((long long)v);
}

// One step of all five iterations for this lane.  MODE 0: every lane and iteration is on an interior point (no masks,
// no boundary coefficients).  MODE 1: interior ROWS, any column -- the first and the last strips of a panel, whose
// column tests do not depend on the step and are hoisted out of the loop by the compiler; these strips head the chain
// every other strip waits for, so their step time is the period of the whole pipeline.  MODE 2: anything (the first
// and last batches of every strip).  P[k]: this lane's result of iteration k+1 in the previous step; A1/A2[k]: partial
// sum after iteration k+1, one / two steps ago.
template <int MODE, int SPC>
__device__ __forceinline__ void gss_step(const GssArgs &a, int xa, int xb, int s, int lane, int colbase, double b1, double g1, double dinv,
                                         double (&P)[5], double (&A1)[4], double (&A2)[4], double *zin_slot, double *zout_slot,
                                         const double *hs_x /*lane 0: exchange slot of this step (x1..x5), or null*/,
                                         const double *hs_a /*lane 0: slot of the previous step (sum1..sum4 at +5)*/,
                                         double *hout /*lane 31: slot s-31, or null*/,
                                         const double *cf = nullptr /*MODE 1: this lane's coefficient table, entry j at cf[32*j]*/,
                                         bool act5 = true /*MODE 1: the column of iteration 5 belongs to the panel*/)
{
    constexpr bool FAST = MODE == 0, TAB = MODE == 1, ROWS_IN = MODE == 1 || MODE == 3;
    const unsigned full = 0xffffffffu;
    double S[5], T[4];
    double OX[5], OA[4];   // what lane 31 hands to the next strip: never GSS_EMPTY (inactive points hand over 0)
#pragma unroll
    for (int k = 0; k < 5; ++k) S[k] = __shfl_up_sync(full, P[k], 1);
#pragma unroll
    for (int k = 0; k < 4; ++k) T[k] = __shfl_up_sync(full, A2[k], 1);
    if (hs_x != nullptr) {   // lane 0 of a strip with a left neighbour
#pragma unroll
        for (int k = 0; k < 5; ++k) S[k] = hs_x[k];
#pragma unroll
        for (int k = 0; k < 4; ++k) T[k] = hs_a[5 + k];
    }
    const double zin = *zin_slot;
#pragma unroll
    for (int k = 5; k >= 1; --k) {
        const int r = s - lane - (k - 1);
        const int ck = colbase - (k - 1);
        const bool row_ok = ROWS_IN || (unsigned)r < (unsigned)a.H;
        const bool act = FAST || (TAB ? act5 : (row_ok && ck >= xa && ck < xb));   // [xa, xb): columns this panel sweeps
        double rhs;
        if (k == 1) {
            rhs = dinv * zin;                                                  // x = dinv*z (:470-481)
        } else {
            const double xr = P[k - 2], xu = S[k - 2];                         // iterate k-1 at (c+1,r) and (c,r+1)
            if (FAST) {
                rhs = b1 * xr + g1 * xu;
            } else if (TAB) {
                rhs = cf[32 * (2 * k - 2)] * xr + g1 * xu;                     // coefficient 0 where the reference has no such term
            } else {                                                           // (D-inverse)*U*x (:486-530)
                const double betU = (ck == 0) ? 2 * b1 : b1;                    // mirror coefficient at the DOMAIN edge only
                const double gamU = (!ROWS_IN && r == 0 && a.bot_bnd) ? 2 * g1 : g1;
                const double t1 = betU * xr, t2 = gamU * xu;
                const bool has_right = ck < xb - 1, lasty = !ROWS_IN && (r == a.H - 1);
                rhs = has_right ? (lasty ? t1 : t1 + t2) : (lasty ? 0.0 : t2);
            }
        }
        double v = rhs;                                                        // [(I - (D-inverse)*L)]-inverse (:532-566)
        if (FAST) {
            v = (v + b1 * S[k - 1]) + g1 * P[k - 1];
        } else if (TAB) {
            v = (v + cf[32 * (2 * k - 1)] * S[k - 1]) + g1 * P[k - 1];
        } else {
            const double betL = (ck == a.mx - 1) ? 2 * b1 : b1;
            const double gamL = (!ROWS_IN && r == a.H - 1 && a.top_bnd) ? 2 * g1 : g1;
            if (ck > xa) v = v + betL * S[k - 1];
            if (ROWS_IN || r > 0) v = v + gamL * P[k - 1];
        }
        const double sum = (k == 1) ? 0.0 + v : T[k - 2] + v;                  // z := z + x (:568-570), z zeroed at :478
        if (k == 5) {
            if (act) *zout_slot = sum;
        } else {
            A2[k - 1] = A1[k - 1];
            A1[k - 1] = sum;
            OA[k - 1] = (FAST || TAB || act) ? sum : 0.0;
        }
        P[k - 1] = v;
        OX[k - 1] = (FAST || TAB || act) ? v : 0.0;
    }
    if (hout != nullptr) {   // lane 31 of a strip with a right neighbour: one 80-byte slot per step, five 16-byte stores
        double2 *h2 = reinterpret_cast<double2 *>(hout);
        __stcg(h2 + 0, make_double2(OX[0], OX[1]));
        __stcg(h2 + 1, make_double2(OX[2], OX[3]));
        __stcg(h2 + 2, make_double2(OX[4], OA[0]));
        __stcg(h2 + 3, make_double2(OA[1], OA[2]));
        __stcg(h2 + 4, make_double2(OA[3], 0.0));
    }
}

// CTA = SPC compute warps (one species each) + ONE I/O warp (a second one is a tuning option).  The compute warps do
// nothing but steps: the strips form a dependency chain (strip I runs >= 40 steps behind strip I-1), so the time of the
// whole sweep is (rows + 48 x strips) x the time of one warp's step, and everything else a compute warp used to do per
// batch -- z loads, finished rows to memory, polling the exchange words (two thirds of its time, ncu r02_gs_stream) --
// sat on that chain.  The I/O warp moves the z rows of all SPC species in and out of the window and polls and stages
// their exchange words while the compute warps run the batch's steps; one barrier per batch.
template <int NS, int SPC>
__global__ void __launch_bounds__(32 * (SPC + 2), 2) gs_stream_kernel(const __grid_constant__ GssArgs a, unsigned *ticket, int total,
                                                                      const double *zin, double *zout, double *__restrict__ hbuf)
{
    constexpr int PITCH = 36 * SPC + ((SPC % 2 == 0) ? 1 : 0);   // (PITCH - SPC) odd: the skewed accesses hit 32 different banks
    constexpr int NHL = (GSS_B * GSS_HS + 31) / 32;              // exchange words per lane, species and batch
    constexpr int HSB = GSS_B * GSS_HS;                          // one staged batch of exchange slots
    extern __shared__ double sm[];
    double *zwin = sm;                                           // [GSS_WIN][PITCH]
    double *hst = sm + GSS_WIN * PITCH;                          // [SPC][3][HSB]: read by this batch / the previous one / being filled
    double *ctab = hst + SPC * 3 * HSB;                          // [SPC][10][32]: edge strips, coefficients per (iteration, lane)
    __shared__ int cur;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const bool io = w >= SPC, zrole = w == SPC;      // zrole: the z-row warp; the other I/O warp handles the exchange words
    const bool one_io = blockDim.x == 32 * (SPC + 1);   // launched with a single I/O warp: it does both jobs in turn
    const int mx = a.mx, H = a.H;
    const i64 mxns = (i64)mx * NS;

    for (;;) {
        __syncthreads();
        if (tid == 0) cur = (int)atomicAdd(ticket, 1u);
        __syncthreads();
        const int item = cur;
        if (item >= total) return;
        // items are numbered strip-major: strip 0 of every panel first, so that a CTA only ever waits for an earlier ticket
        const int I = item / (a.npanels * a.ng);                 // strip within its panel
        const int pg = item - I * (a.npanels * a.ng);
        const int pn = pg / a.ng, g = pg - pn * a.ng, g0 = g * SPC;
        const int c0 = pn * a.pw, c1 = min(c0 + a.pw, mx);       // columns this panel owns (stores)
        const int xa = max(c0 - a.ovl, 0), xb = min(c1 + a.ovr, mx);   // columns it sweeps
        const int nst = (xb - xa + 4 + 31) / 32;                 // strips of this panel (the last one may be narrower than the others')
        const bool has_prod = I > 0, has_cons = I < nst - 1;
        if (I >= nst) continue;                                  // a narrower last panel has fewer strips
        const i64 hstride = a.hp * GSS_HS;                       // exchange words per (strip boundary, species)
        double *hprod0 = has_prod ? hbuf + (((i64)pn * a.nstrips + I - 1) * NS + g0) * hstride : nullptr;

        if (io) {
            // ================= I/O warp: lane-flat index f = m*32 + lane over the 32 x SPC (point, species) pairs of a row
            int fpt[SPC];
            i64 offl[SPC], offs[SPC];
            bool cl_ok[SPC], cs_ok[SPC];
#pragma unroll
            for (int m = 0; m < SPC; ++m) {
                const int f = m * 32 + lane, pt = f / SPC, si = f - pt * SPC;
                const int cl = xa + 32 * I + pt, cs = xa + 32 * I - 4 + pt;   // columns loaded / stored
                fpt[m] = f;
                cl_ok[m] = cl < xb;
                cs_ok[m] = cs >= c0 && cs < c1;
                offl[m] = (i64)cl * NS + g0 + si;
                offs[m] = (i64)cs * NS + g0 + si;
            }
            // z rows travel one batch ahead of their use: rows [s0+16, s0+24) are requested during batch s0, enter the
            // window during batch s0+8 and are first read by the steps of batch s0+16 -- a whole batch of steps to cover
            // the DRAM latency.  Before the loop: rows [0, 8).
            double zl[GSS_B][SPC];
            auto load_rows = [&](int r0) {
                const bool whole = r0 + GSS_B <= H;
#pragma unroll
                for (int j = 0; j < GSS_B; ++j)
#pragma unroll
                    for (int m = 0; m < SPC; ++m)
                        zl[j][m] = (cl_ok[m] && (whole || r0 + j < H)) ? GSS_LDZ(zin + (i64)(r0 + j) * mxns + offl[m]) : 0.0;
            };
            if (zrole) load_rows(0);
            for (int s0 = -GSS_B; s0 <= a.s0_last; s0 += GSS_B) {
                const int bw = ((s0 / GSS_B) + 2) % 3;             // staging buffer filled now, read by the next batch
                __syncthreads();
                if (zrole) {
                    // ---- rows [s0+8, s0+16) enter the window (read from the next batch on); request the rows after them ----
#pragma unroll
                    for (int j = 0; j < GSS_B; ++j) {
                        double *wdst = zwin + GSS_ROW(s0 + GSS_B + j) * PITCH + 4 * SPC;
#pragma unroll
                        for (int m = 0; m < SPC; ++m) wdst[fpt[m]] = zl[j][m];
                    }
                    load_rows(s0 + 2 * GSS_B);
                    // ---- finished rows [s0-36-B, s0-36) leave the window (B = 8: [s0-44, s0-36)) ----
                    if (s0 >= 37) {
                        const int r0 = s0 - 36 - GSS_B;
#pragma unroll
                        for (int j = 0; j < GSS_B; ++j) {
                            const int r = r0 + j;
                            if (r >= 0 && r < H) {
                                const double *wsrc = zwin + GSS_ROW(r) * PITCH;
                                double t[SPC];
#pragma unroll
                                for (int m = 0; m < SPC; ++m) t[m] = wsrc[fpt[m]];
#pragma unroll
                                for (int m = 0; m < SPC; ++m)
                                    if (cs_ok[m]) zout[(i64)r * mxns + offs[m]] = t[m];
                            }
                        }
                    }
                    if (!one_io) continue;
                }
                if (has_prod) {
                    // Every exchange word is its own flag: the buffer holds GSS_EMPTY (a NaN no arithmetic produces) until the
                    // left strip's lane 31 stores the value (one aligned 8-byte store), and goes back to GSS_EMPTY as soon as it
                    // has been taken.  No fences, no progress counters: a word is either there or not.  All SPC x NHL words
                    // of this lane are polled together, round after round.
                    double hl[SPC][NHL];
                    bool need[SPC][NHL];
                    bool pending = false;
#pragma unroll
                    for (int q = 0; q < SPC; ++q)
#pragma unroll
                        for (int m = 0; m < NHL; ++m) {
                            const int idx = lane + 32 * m;
                            need[q][m] = idx < HSB && s0 + GSS_B + idx / GSS_HS <= H + 3;
                            hl[q][m] = __longlong_as_double(GSS_EMPTY);
                            pending = pending || need[q][m];
                        }
                    while (pending) {
#pragma unroll
                        for (int q = 0; q < SPC; ++q)
#pragma unroll
                            for (int m = 0; m < NHL; ++m)
                                if (need[q][m] && __double_as_longlong(hl[q][m]) == GSS_EMPTY)
                                    hl[q][m] = gss_poll(hprod0 + q * hstride + (i64)(s0 + GSS_B) * GSS_HS + lane + 32 * m);
                        pending = false;
#pragma unroll
                        for (int q = 0; q < SPC; ++q)
#pragma unroll
                            for (int m = 0; m < NHL; ++m)
                                pending = pending || (need[q][m] && __double_as_longlong(hl[q][m]) == GSS_EMPTY);
                        if (pending) __nanosleep(a.poll_ns);
                    }
#pragma unroll
                    for (int q = 0; q < SPC; ++q)
#pragma unroll
                        for (int m = 0; m < NHL; ++m)
                            if (need[q][m]) {
                                const int idx = lane + 32 * m;
                                __stcg(hprod0 + q * hstride + (i64)(s0 + GSS_B) * GSS_HS + idx, __longlong_as_double(GSS_EMPTY));
                                hst[(q * 3 + bw) * HSB + idx] = hl[q][m];
                            }
                }
            }
            continue;
        }

        // ================= compute warp w: species g0 + w
        const int sp = g0 + w;
        const double b1 = a.beta1[sp], g1 = a.gamma1[sp], dinv = a.dinv[sp];
        const int colbase = xa + 32 * I + lane;
        double *hcons = has_cons ? hbuf + (((i64)pn * a.nstrips + I) * NS + sp) * hstride : nullptr;
        const double *hmine = hst + w * (3 * HSB);
        // interior strip: every column of every iteration has both neighbours
        const bool strip_fast = (I >= 1) && (xa + 32 * I + 31 <= xb - 2);
        // Edge strips (the first one and the last ones of a panel) away from the first / last rows run the interior step
        // with per-(iteration, lane) coefficients from a table: b1, 2*b1 at a mirrored domain edge, or 0 where the reference
        // has no left / right term at all (x + 0*y == x up to the sign of a zero).  These strips head the dependency chain,
        // so their step time is everybody's (the select-based step, MODE 3, has 1.6 times the instructions).  The table
        // form hands lane 31's values on unmasked, so it needs lane 31 inside the panel whenever there is a consumer.
        const bool edge_tab = !strip_fast && a.edge_fast && (!has_cons || xa + 32 * I + 27 < xb);
        double *cf = ctab + w * 320 + lane;
        bool act5 = true;
        if (edge_tab) {
#pragma unroll
            for (int k = 1; k <= 5; ++k) {
                const int ck = colbase - (k - 1);
                cf[32 * (2 * k - 2)] = (ck < xb - 1) ? ((ck == 0) ? 2 * b1 : b1) : 0.0;
                cf[32 * (2 * k - 1)] = (ck > xa) ? ((ck == a.mx - 1) ? 2 * b1 : b1) : 0.0;
            }
            act5 = colbase - 4 >= xa && colbase - 4 < xb;
            __syncwarp();
        }
        double P[5] = {0, 0, 0, 0, 0}, A1[4] = {0, 0, 0, 0}, A2[4] = {0, 0, 0, 0};

        for (int s0 = -GSS_B; s0 <= a.s0_last; s0 += GSS_B) {
            const int br = ((s0 / GSS_B) + 1) % 3, bp = ((s0 / GSS_B) + 3) % 3;   // staging buffers of this batch / the previous one
            __syncthreads();
            if (s0 >= 0 && s0 <= H + 35) {
                const bool rows_in = (s0 - 35 >= 1) && (s0 + GSS_B - 1 <= H - 2);   // rows 1 .. H-2 for every lane and iteration
                const bool fast = strip_fast && rows_in;
                const double *hsb = (has_prod && lane == 0) ? hmine + br * HSB : nullptr;
                const double *hsp = hmine + bp * HSB + (GSS_B - 1) * GSS_HS;       // last slot of the previous batch
                double *hcl = (has_cons && lane == 31) ? hcons : nullptr;
                // (not unrolled beyond 2: eight inlined steps are ~28 KB of code per path and the warps of an SM sit at
                // different places in it)
                if (fast) {
#pragma unroll 2
                    for (int j = 0; j < GSS_B; ++j) {
                        const int s = s0 + j;
                        double *zi = zwin + GSS_ROW(s - lane) * PITCH + (lane + 4) * SPC + w;
                        double *zo = zwin + GSS_ROW(s - lane - 4) * PITCH + lane * SPC + w;
                        gss_step<0, SPC>(a, xa, xb, s, lane, colbase, b1, g1, dinv, P, A1, A2, zi, zo, hsb ? hsb + j * GSS_HS : nullptr,
                                         j ? hsb + (j - 1) * GSS_HS : hsp, hcl ? hcl + (i64)(s - 31) * GSS_HS : nullptr);
                    }
                } else if (rows_in && edge_tab) {      // an edge strip away from the first and last rows (every slot of the batch is awaited)
#pragma unroll 2
                    for (int j = 0; j < GSS_B; ++j) {
                        const int s = s0 + j;
                        double *zi = zwin + GSS_ROW(s - lane) * PITCH + (lane + 4) * SPC + w;
                        double *zo = zwin + GSS_ROW(s - lane - 4) * PITCH + lane * SPC + w;
                        gss_step<1, SPC>(a, xa, xb, s, lane, colbase, b1, g1, dinv, P, A1, A2, zi, zo, hsb ? hsb + j * GSS_HS : nullptr,
                                         j ? hsb + (j - 1) * GSS_HS : hsp, hcl ? hcl + (i64)(s - 31) * GSS_HS : nullptr, cf, act5);
                    }
                } else if (rows_in && a.edge_fast) {   // same with selects: lane 31 may lie outside the panel
#pragma unroll 1
                    for (int j = 0; j < GSS_B; ++j) {
                        const int s = s0 + j;
                        double *zi = zwin + GSS_ROW(s - lane) * PITCH + (lane + 4) * SPC + w;
                        double *zo = zwin + GSS_ROW(s - lane - 4) * PITCH + lane * SPC + w;
                        gss_step<3, SPC>(a, xa, xb, s, lane, colbase, b1, g1, dinv, P, A1, A2, zi, zo, hsb ? hsb + j * GSS_HS : nullptr,
                                         j ? hsb + (j - 1) * GSS_HS : hsp, hcl ? hcl + (i64)(s - 31) * GSS_HS : nullptr);
                    }
                } else {
#pragma unroll 1
                    for (int j = 0; j < GSS_B; ++j) {
                        const int s = s0 + j;
                        double *zi = zwin + GSS_ROW(s - lane) * PITCH + (lane + 4) * SPC + w;
                        double *zo = zwin + GSS_ROW(s - lane - 4) * PITCH + lane * SPC + w;
                        // only the slots the right-hand strip waits for (steps it has an active lane 0 in) are written
                        gss_step<2, SPC>(a, xa, xb, s, lane, colbase, b1, g1, dinv, P, A1, A2, zi, zo, hsb ? hsb + j * GSS_HS : nullptr,
                                         j ? hsb + (j - 1) * GSS_HS : hsp,
                                         (hcl && s >= 31 && s - 31 <= H + 3) ? hcl + (i64)(s - 31) * GSS_HS : nullptr);
                    }
                }
            }
        }
    }
}

static int gss_spc(int ns) { return ns % 4 == 0 ? 4 : ns % 5 == 0 ? 5 : ns % 3 == 0 ? 3 : ns % 2 == 0 ? 2 : 1; }

template <int NS, int SPC>
static int gss_launch(const GssArgs &a, unsigned *ticket, int total, const double *zin, double *zout, double *hbuf)
{
    DkbCtx *c = g_ctx;
    constexpr int PITCH = 36 * SPC + ((SPC % 2 == 0) ? 1 : 0);
    const size_t smem = sizeof(double) * ((size_t)GSS_WIN * PITCH + (size_t)SPC * 3 * GSS_B * GSS_HS + (size_t)SPC * 320);
    static int per_sm = 0;
    if (!per_sm) {
        cudaFuncSetAttribute(gs_stream_kernel<NS, SPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(gs_stream_kernel<NS, SPC>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gs_stream_kernel<NS, SPC>, 32 * (SPC + 2), smem);
        per_sm = std::max(1, per_sm);
    }
    // Two CTAs per SM: the strips form one dependency chain (strip I runs >= 40 steps behind strip I-1), so the time
    // is (rows + 48 x strips) x the time of ONE warp's step, and a third CTA per SM slows every warp's step by more than
    // the extra tickets in flight gain (measured 4096^2: 6.3 ms with two, 7.6 ms with three).  debug_variant 31 / 33:
    // tuning runs with one / three.
    const int cap = (c->debug_variant >= 31 && c->debug_variant <= 33) ? std::min(per_sm, c->debug_variant - 30) : std::min(per_sm, 2);
    const int grid = std::min(total, c->nsm * cap);
    // one I/O warp does both jobs; a second one (debug_variant 42: z rows / exchange words split) shortens the I/O warp's
    // own batch but slows the compute warps it shares schedulers with by more: 6.05 against 5.65 ms at 4096^2
    const int nio = c->debug_variant == 42 ? 2 : 1;
    gs_stream_kernel<NS, SPC><<<grid, 32 * (SPC + nio), smem, c->stream>>>(a, ticket, total, zin, zout, hbuf);
    return 0;
}

// zout := (approximately) A_S^-1 zin over `my` rows (zout == zin: in place; required when npanels == 1 is not... see below).
// npanels > 1: the columns are cut into panels that are swept independently, each extended by `ovl` columns of its left
// and `ovr` columns of its right neighbour -- the overlapped decomposition the slabs already use in y (gs.cu), for the
// same reason: the triangular solve couples a point to the columns left of it through paths of total weight
// <= (beta/(1-gamma))^m over m columns (at most (1/2)^m), the U step reaches one column to the right per iteration.  That
// shortens the strip pipeline from mx/32 to pw/32 stages, which is what a short, wide slab (8 GPUs) needs.  Panels read
// columns their neighbours own, so zin and zout must then be different arrays.
int gs_stream_sweep(double hl0, int my, bool bot_bnd, bool top_bnd, const double *zin, double *zout, int npanels)
{
    DkbCtx *c = g_ctx;
    const WebPar &wp = c->web;
    const int ns = wp.ns;
    GssArgs a;
    a.mx = wp.mx;
    a.H = my;
    a.bot_bnd = bot_bnd;
    a.top_bnd = top_bnd;
    a.edge_fast = getenv("DKB_GS_NOEDGE") ? 0 : 1;
    a.poll_ns = getenv("DKB_GS_POLL_NS") ? (unsigned)atoi(getenv("DKB_GS_POLL_NS")) : 20u;
    for (int i = 0; i < ns; ++i) {   // :447-457
        const double elamda = 1.0 / (1.0 + 2 * hl0 * (wp.cox[i] + wp.coy[i]));
        a.beta1[i] = hl0 * wp.cox[i] * elamda;
        a.gamma1[i] = hl0 * wp.coy[i] * elamda;
        a.dinv[i] = elamda;
    }
    const int spc = gss_spc(ns);
    a.ng = ns / spc;
    if (npanels > 1 && zin == zout) return dkb_fail(DKB_E_ARG, "Gauss-Seidel factor: column panels need separate in / out arrays");
    npanels = std::max(1, std::min(npanels, (wp.mx + 63) / 64));
    a.pw = (((wp.mx + npanels - 1) / npanels + 31) / 32) * 32;
    a.npanels = (wp.mx + a.pw - 1) / a.pw;
    a.ovl = a.npanels > 1 ? 32 : 0;
    a.ovr = a.npanels > 1 ? 8 : 0;
    // iteration k covers columns shifted k-1 to the left: 4 more columns at the right end
    a.nstrips = (std::min(a.pw + a.ovl + a.ovr, wp.mx) + 4 + 31) / 32;
    a.hp = ((i64)(my + 64 + GSS_B - 1) / GSS_B) * GSS_B + 2 * GSS_B;
    a.s0_last = ((my + 36 + GSS_B - 1) / GSS_B) * GSS_B;
    const size_t hwords = (size_t)a.npanels * a.nstrips * ns * a.hp * GSS_HS;
    if (hwords > c->gss_hwords) {
        cudaFree(c->gss_hbuf);
        c->gss_hbuf = nullptr;
        c->gss_hwords = 0;
        if (cudaMalloc(&c->gss_hbuf, sizeof(double) * hwords) != cudaSuccess)
            return dkb_fail(DKB_E_CUDA, "Gauss-Seidel factor: cannot allocate the strip exchange buffer (%zu MB)", hwords >> 17);
        // every word GSS_EMPTY; the kernel leaves the buffer in that state (each word is taken exactly once)
        cudaMemsetAsync(c->gss_hbuf, 0xFF, sizeof(double) * hwords, c->stream);
        c->gss_hwords = hwords;
    }
    if (!c->gss_ticket && cudaMalloc(&c->gss_ticket, sizeof(unsigned)) != cudaSuccess)
        return dkb_fail(DKB_E_CUDA, "Gauss-Seidel factor: cannot allocate the ticket counter");
    cudaMemsetAsync(c->gss_ticket, 0, sizeof(unsigned), c->stream);
    const int total = a.nstrips * a.npanels * a.ng;
    int rc = 0;
#define GSS_CASE(n, s) case n: rc = gss_launch<n, s>(a, c->gss_ticket, total, zin, zout, c->gss_hbuf); break;
    switch (ns) {
        GSS_CASE(1, 1) GSS_CASE(2, 2) GSS_CASE(3, 3) GSS_CASE(4, 4) GSS_CASE(5, 5) GSS_CASE(6, 3) GSS_CASE(7, 1) GSS_CASE(8, 4)
        GSS_CASE(9, 3) GSS_CASE(10, 5) GSS_CASE(11, 1) GSS_CASE(12, 4) GSS_CASE(13, 1) GSS_CASE(14, 2) GSS_CASE(15, 5)
        GSS_CASE(16, 4) GSS_CASE(17, 1) GSS_CASE(18, 3) GSS_CASE(19, 1) GSS_CASE(20, 4) GSS_CASE(21, 3) GSS_CASE(22, 2)
        GSS_CASE(23, 1) GSS_CASE(24, 4) GSS_CASE(25, 5) GSS_CASE(26, 2) GSS_CASE(27, 3) GSS_CASE(28, 4) GSS_CASE(29, 1)
        GSS_CASE(30, 5) GSS_CASE(31, 1) GSS_CASE(32, 4)
    default: return dkb_fail(DKB_E_UNSUPPORTED, "Gauss-Seidel factor: ns=%d out of range 1..32", ns);
    }
#undef GSS_CASE
    COUNT_LAUNCH();
    dkb_cuda_check("gs_stream_sweep");
    return rc;
}

#ifdef GSS_DEBUG
extern "C" void dkb_debug_gss_clocks(long long *out, int n)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, gss_dbg, sizeof(long long) * n);
}
#endif
