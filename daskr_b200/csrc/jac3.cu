// jac3.cu -- K5: jac_rbdpre (src/daskr_rbdpre.f90:158-251) + dgefa (src/linpack.f90:325-429) for the food-web
// problem, ONE thread mapping for both phases: a lane owns TWO rows (l and l + ns/2) of one mesh point's ns x ns
// block, ns/2 lanes make a point, floor(32 / (ns/2)) points share a warp and run in lockstep (ns = 20: 3 points on
// 30 lanes).  The rows never leave their lane's registers between the difference quotients and the factorisation.
//
// Against jac2.cu (thread per (point, column) for the FD phase, block staged in shared memory, one warp of 20 live
// lanes per block for the LU phase; ncu: 3970 warp instructions per point, 3013 of them in the LU phase):
//   * FD phase, row k of R at the state perturbed in species js (rates, example_web.f90:290-297):
//         s_js = (..((c_0 a_0k + c_1 a_1k) + ..) + (c_js+del_js) a_js,k) + c_js+1 a_js+1,k) + .. + c_ns-1 a_ns-1,k
//     the products m_j = c_j a_jk do not depend on js and the partial sum over j < js is the same running prefix for
//     every js, so a row costs ns products + ns(ns+1)/2 additions instead of ns^2 of each -- every sum still adds
//     the same terms in the same order, so the quotients are bit-identical (tests/test_gpu_parity.py);
//   * LU phase: one shuffle per column broadcasts the pivot row's tail to all points of the warp at once (every
//     collective names the full warp, each lane its own source lane: no per-group masks, which the compiler serialises);
//     the pivot row's entries are final at that moment and are written to memory right away, so from then on the
//     registers of finished rows are dead and the row updates run UNCONDITIONALLY on both rows of every lane
//     (multiply + add per element, no selects).  Logical row interchanges as in lu_device.cuh: data never moves,
//     LINPACK's row numbers do.
//   * a zero pivot (linpack.f90:394-397 skips the elimination step) cannot be honoured by unconditional updates: the
//     kernel only DETECTS it (first_flag) and the caller recomputes with the exact kernels of jac2.cu / web.cu.  A
//     singular block is an error return of DASKR's jac (ierr /= 0) either way, so that path is rare.
#include "web_dev.cuh"
#include "lu_device.cuh"

#define J3_WARPS 4

template <int NS, int MINB>
__global__ void __launch_bounds__(32 * J3_WARPS, MINB)
web_jac3_kernel(WebDev w, int nsd, const double *__restrict__ c, const double *__restrict__ rewt, double cj,
                double *__restrict__ bd, int *__restrict__ ipbd, int *first_flag)
{
    constexpr int L = NS / 2;            // lanes per point
    constexpr int BPW = 32 / L;          // points per warp
    constexpr int NSQ = NS * NS;
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(NS % 2 == 0 && L <= 16 && BPW <= 3, "two rows per lane, at most three points per warp");
    __shared__ double s_ac[NSQ];                         // acoef, (k,j) at [j*NS+k]
    __shared__ double s_c[J3_WARPS][BPW + 1][NS];        // the point's concentrations (slot BPW: the spare lanes)
    __shared__ double s_up[J3_WARPS][BPW + 1][NS];       // c_js + del_js         (jac_rbdpre:223)
    __shared__ double s_fd[J3_WARPS][BPW + 1][NS];       // -1/del_js             (:224)
    const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
    for (int i = tid; i < NSQ; i += 32 * J3_WARPS) s_ac[i] = __ldg(w.acoef + i);
    __syncthreads();
    // Lanes past the last whole point (ns = 20: lanes 30, 31) run along as a point of their own whose results are
    // never stored, so that every shuffle, vote and reduction below is a full-warp one.
    const bool spare = lane >= BPW * L;
    const int g = spare ? BPW : lane / L, l = spare ? lane - BPW * L : lane - g * L;
    const unsigned gmask = spare ? ~((1u << (BPW * L)) - 1u) : (((1u << L) - 1u) << (g * L));
    const i64 npts = (i64)w.mx * w.my;
    const i64 ntrip = (npts + BPW - 1) / BPW;
    const double bcoA = __ldg(w.bcoef + l), bcoB = __ldg(w.bcoef + l + L);
    double *sc = s_c[wp][g], *sup = s_up[wp][g], *sfd = s_fd[wp][g];

    // The factors leave through a staging buffer in shared memory and are written by the whole CTA, PTS consecutive mesh
    // points (whole 32-byte sectors) per matrix element.  Stored straight from the LU loop -- 8 bytes per point and
    // store, 24 bytes per warp -- every store made the L2 fetch its sector first once the array outgrew what the
    // cache merges: ncu at 4096^2 x 20: 65 GB read and 99 GB written for 55 GB of blocks, 84 ms where 16 x the 1024^2
    // time is 60.
    constexpr int PTS = BPW * J3_WARPS, SP = PTS + 1;   // pitch PTS + 1: the rows of a point fall into different banks
    extern __shared__ double j3_stage[];                // [NSQ][SP] doubles, then [NS][SP] ints
    int *j3_piv = reinterpret_cast<int *>(j3_stage + NSQ * SP);
    const int pidx = wp * BPW + (spare ? 0 : g);
    double *gb = j3_stage + pidx;                       // element e of this lane's point at gb[e * SP]

    for (i64 tb = (i64)blockIdx.x * J3_WARPS; tb < ntrip; tb += (i64)gridDim.x * J3_WARPS) {
        const i64 trip = tb + wp;
        const i64 p = trip * BPW + (spare ? 0 : g);
        const bool live = !spare && p < npts;            // (a warp past the last group just keeps the CTA's barriers)
        const i64 pp = p < npts ? p : npts - 1;
        // ---- the point's state, increments and -1/del, two species per lane
        {
            const double cA = __ldg(c + pp * NS + l), cB = __ldg(c + pp * NS + l + L);
            const double rA = __ldg(rewt + pp * NS + l), rB = __ldg(rewt + pp * NS + l + L);
            const double dA = fmax(w.srur * fabs(cA), 1e-2 / rA);   // jac_rbdpre:222
            const double dB = fmax(w.srur * fabs(cB), 1e-2 / rB);
            __syncwarp();                                            // the previous point's readers are done
            sc[l] = cA;
            sc[l + L] = cB;
            sup[l] = cA + dA;
            sup[l + L] = cB + dB;
            sfd[l] = -1.0 / dA;
            sfd[l + L] = -1.0 / dB;
            __syncwarp();
        }
        const int jyl = (int)(pp / w.mx);
        const int jx = (int)(pp - (i64)jyl * w.mx);
        const int jyg = w.jy0 + jyl;
        const double yy = jyg * w.dy, xx = jx * w.dx;
        const double fac = 1.0 + w.alpha * xx * yy + w.beta * __ldg(w.sx + jx) * __ldg(w.sy + jyg);
        const double bfA = bcoA * fac, bfB = bcoB * fac;
        const double cA = sc[l], cB = sc[l + L];

        // ---- phase FD: rows l and l + L of the block
        double a0[NS], a1[NS];
        {
            double mA[NS], mB[NS];
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                const double cv = sc[j];
                mA[j] = cv * s_ac[j * NS + l];
                mB[j] = cv * s_ac[j * NS + l + L];
            }
            double rA = 0.0, rB = 0.0;
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                rA = rA + mA[j];
                rB = rB + mB[j];
            }
            const double r0A = cA * (bfA + rA), r0B = cB * (bfB + rB);   // R_k(u), what res left in rpar (f:250)
            double preA = 0.0, preB = 0.0;
#pragma unroll
            for (int js = 0; js < NS; ++js) {
                const double up = sup[js], fd = sfd[js];
                double sA = preA + up * s_ac[js * NS + l];
                double sB = preB + up * s_ac[js * NS + l + L];
#pragma unroll
                for (int j = js + 1; j < NS; ++j) {
                    sA = sA + mA[j];
                    sB = sB + mB[j];
                }
                const double r1A = ((js == l) ? up : cA) * (bfA + sA);
                const double r1B = ((js == l + L) ? up : cB) * (bfB + sB);
                double vA = (r1A - r0A) * fd, vB = (r1B - r0B) * fd;      // (:226-228)
                // + cj*I_d (:240-244); x + (-0.0) == x for every x, so the off-diagonal lanes add nothing
                vA = vA + ((js == l && js < nsd) ? cj : -0.0);
                vB = vB + ((js == l + L && js < nsd) ? cj : -0.0);
                a0[js] = vA;
                a1[js] = vB;
                preA = preA + mA[js];
                preB = preB + mB[js];
            }
        }

        // ---- phase LU: dgefa with logical row interchanges, two rows per lane
        int r0 = l, r1 = l + L;          // LINPACK row currently held in a0 / a1 (columns >= k)
        int ipvA = NS, ipvB = NS;        // ipvt(l+1), ipvt(l+L+1)
        bool zero_pivot = false;
#pragma unroll
        for (int k = 0; k < NS - 1; ++k) {
            // idamax: first maximum of |a(i,k)| over the rows i >= k, on the integer image of the magnitudes.
            // Each lane first settles its own two rows, then one max / max / min reduction per point of the warp.
            const unsigned long long k0 = r0 >= k ? (unsigned long long)__double_as_longlong(fabs(a0[k])) : 0ull;
            const unsigned long long k1 = r1 >= k ? (unsigned long long)__double_as_longlong(fabs(a1[k])) : 0ull;
            const unsigned q0 = r0 >= k ? (unsigned)r0 : 0xffu, q1 = r1 >= k ? (unsigned)r1 : 0xffu;
            const bool pickB = k1 > k0 || (k1 == k0 && q1 < q0);
            const unsigned long long km = pickB ? k1 : k0;
            const unsigned qm = pickB ? q1 : q0;
            const unsigned hi = (unsigned)(km >> 32), lo = (unsigned)km;
            // first the high words (sign-free exponent + 20 mantissa bits): almost always that maximum is held by one
            // lane of the point and the search is over; ties in the high word fall through to the exact comparison
            unsigned mhi = 0;
#pragma unroll
            for (int gg = 0; gg < BPW; ++gg) {
                const unsigned m = __reduce_max_sync(FULL, g == gg ? hi : 0u);
                if (g == gg) mhi = m;
            }
            const unsigned tied = __ballot_sync(FULL, hi == mhi) & gmask;
            int lp, pl;                                               // LINPACK's l (0-based); lane holding that row
            bool isP;
            if (__all_sync(FULL, spare || (tied & (tied - 1)) == 0)) {
                pl = __ffs(tied) - 1;
                isP = lane == pl;
                lp = (int)__shfl_sync(FULL, qm, pl);
            } else {
                lp = 0;
#pragma unroll
                for (int gg = 0; gg < BPW; ++gg) {
                    const bool mine = g == gg && hi == mhi;
                    const unsigned mlo = __reduce_max_sync(FULL, mine ? lo : 0u);
                    const unsigned ml = __reduce_min_sync(FULL, (mine && lo == mlo) ? qm : 0xffu);
                    if (g == gg) lp = (int)ml;
                }
                isP = !spare && (r0 == lp || r1 == lp);
                pl = __ffs(__ballot_sync(FULL, isP) & gmask) - 1;
            }
            const bool isB = isP && r1 == lp;                         // the pivot row sits in a1
            if (k < L) { if (l == k) ipvA = lp + 1; } else { if (l == k - L) ipvB = lp + 1; }
            // the pivot row, columns k..ns-1, is final: row k of U goes to memory from the lane that holds it and to
            // every lane of the point by one shuffle per column
            double tj[NS];
#pragma unroll
            for (int j = k; j < NS; ++j) {
                const double src = isB ? a1[j] : a0[j];
                if (isP && live) gb[(j * NS + k) * SP] = src;
                tj[j] = __shfl_sync(FULL, src, pl);
            }
            const double piv = tj[k];
            zero_pivot = zero_pivot || piv == 0.0;                    // see the header: detected here, redone by the caller
            if (r0 == lp) r0 = k; else if (r0 == k) r0 = lp;          // interchange (logical)
            if (r1 == lp) r1 = k; else if (r1 == k) r1 = lp;
            const double t = -1.0 / piv;
            a0[k] = a0[k] * t;                                        // multipliers (negated); garbage on finished rows
            a1[k] = a1[k] * t;
            if (live && r0 > k) gb[(k * NS + r0) * SP] = a0[k];
            if (live && r1 > k) gb[(k * NS + r1) * SP] = a1[k];
#pragma unroll
            for (int j = k + 1; j < NS; ++j) {
                a0[j] = a0[j] + tj[j] * a0[k];
                a1[j] = a1[j] + tj[j] * a1[k];
            }
        }
        {
            const bool lastB = r1 == NS - 1, lastP = lastB || r0 == NS - 1;
            const double ann = lastB ? a1[NS - 1] : a0[NS - 1];
            if (lastP && live) gb[((NS - 1) * NS + NS - 1) * SP] = ann;
            zero_pivot = zero_pivot || (lastP && ann == 0.0);
        }
        if (live) {
            j3_piv[l * SP + pidx] = ipvA;
            j3_piv[(l + L) * SP + pidx] = ipvB;
            if (zero_pivot) atomicMin(first_flag, (int)(p < 2147483646 ? p + 1 : 2147483646));
        }
        // ---- the CTA's PTS points leave: consecutive threads write consecutive points of one matrix element
        __syncthreads();
        {
            const i64 p0 = tb * BPW;
            for (int idx = tid; idx < NSQ * PTS; idx += 32 * J3_WARPS) {
                const int e = idx / PTS, q = idx - e * PTS;
                const i64 pt = p0 + q;
                if (pt < npts) bd[((pt >> 5) * NSQ + e) * 32 + (pt & 31)] = j3_stage[e * SP + q];
            }
            for (int idx = tid; idx < NS * PTS; idx += 32 * J3_WARPS) {
                const int e = idx / PTS, q = idx - e * PTS;
                const i64 pt = p0 + q;
                if (pt < npts) ipbd[((pt >> 5) * NS + e) * 32 + (pt & 31)] = j3_piv[e * SP + q];
            }
        }
        __syncthreads();
    }
}

template <int NS, int MINB = 3>
static void jac3_launch(const WebDev &w, int nsd, const double *c, const double *rewt, double cj, double *bd, int *ipbd,
                        int *d_first_flag)
{
    constexpr int BPW = 32 / (NS / 2);
    constexpr int SP = BPW * J3_WARPS + 1;
    const size_t smem = sizeof(double) * NS * NS * SP + sizeof(int) * NS * SP;
    static int per_sm = 0;
    if (!per_sm) {
        cudaFuncSetAttribute(web_jac3_kernel<NS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, web_jac3_kernel<NS, MINB>, 32 * J3_WARPS, smem);
        per_sm = std::max(1, per_sm);
    }
    const i64 npts = (i64)w.mx * w.my;
    const i64 nblk = (npts + (i64)BPW * J3_WARPS - 1) / (BPW * J3_WARPS);
    // grid-stride over point groups: consecutive warps take consecutive groups, so the 32-point tiles of the
    // interleaved layout are completed by warps that run at the same time (the partial sectors meet in L2)
    const int grid = (int)std::min<i64>(nblk, (i64)g_ctx->nsm * per_sm * 8);
    web_jac3_kernel<NS, MINB><<<grid, 32 * J3_WARPS, smem, g_ctx->stream>>>(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag);
    g_ctx->launches++;
}

// Returns false if ns is not instantiated.  On return *d_first_flag < INT_MAX means "some block hit a zero pivot":
// the blocks written are then NOT LINPACK's and the caller must run the exact kernel (web_jac_launch does).
bool web_jac3(const WebDev &w, int nsd, const double *c, const double *rewt, double cj, double *bd, int *ipbd,
              int *d_first_flag)
{
    switch (w.ns) {
    case 20: jac3_launch<20>(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag); return true;   // (128 registers / 16 warps per SM: 4.27 ms against 3.75)
    case 22: jac3_launch<22>(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag); return true;
    case 24: jac3_launch<24>(w, nsd, c, rewt, cj, bd, ipbd, d_first_flag); return true;
    default: return false;
    }
}
