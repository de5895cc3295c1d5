// lu_device.cuh -- in-register LINPACK dgefa / dgesl for one ns x ns block per sub-warp group.
// Shared by lu.cu (batched micro-API, psol) and the problem kernels (web.cu, heat.cu) that
// fuse the block solve into the residual sweep.  See lu.cu for the layout and the mapping.
#pragma once
#include "dkb_internal.h"

template <int NS> struct GroupOf { static constexpr int G = NS <= 1 ? 1 : NS <= 2 ? 2 : NS <= 4 ? 4 : NS <= 8 ? 8 : NS <= 16 ? 16 : 32; };

__device__ __forceinline__ double shfl_d(unsigned mask, double v, int src, int width)
{
    return __shfl_sync(mask, v, src, width);
}

// Lanes of this thread's group.  Several groups share a warp when G < 32 and may diverge from
// one another (a zero pivot makes one group skip an elimination step), so every shuffle names
// only its own group's lanes.
template <int G>
__device__ __forceinline__ unsigned group_mask()
{
    if (G >= 32) return 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    return ((1u << (G & 31)) - 1u) << (lane & ~(unsigned)(G - 1));
}

// ---- in-register LU of one block per group; shared by dgefa_batched and the jac kernels ----
// a[j] = A(i,j) for this lane's row i (lanes >= NS carry zeros and never win a pivot search).
// Returns LINPACK's info (uniform over the group); ipv = this lane's ipvt(i+1) (1-based).
template <int NS>
__device__ __forceinline__ int group_dgefa(double (&a)[NS], int li, int &ipv)
{
    constexpr int G = GroupOf<NS>::G;
    const unsigned mask = group_mask<G>();
    int info = 0;
    ipv = NS;   // ipvt(n) = n; overwritten for rows k < n below
#pragma unroll
    for (int k = 0; k < NS - 1; ++k) {
        // idamax over rows k..NS-1 of column k: first maximum of |a|
        double av = (li >= k && li < NS) ? fabs(a[k]) : -1.0;
        int ai = li;
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            double bv = __shfl_xor_sync(mask, av, o, G);
            int bi = __shfl_xor_sync(mask, ai, o, G);
            if (bv > av || (bv == av && bi < ai)) {
                av = bv;
                ai = bi;
            }
        }
        const int l = ai;                 // uniform in the group
        if (li == k) ipv = l + 1;
        const double piv = shfl_d(mask, a[k], l, G);
        if (piv == 0.0) {                 // zero pivot: column already triangularised
            info = k + 1;
            continue;
        }
        // interchange a(l,k) <-> a(k,k)
        const double akk = shfl_d(mask, a[k], k, G);
        if (l != k) {
            if (li == l) a[k] = akk;
            if (li == k) a[k] = piv;
        }
        // multipliers (negated)
        const double t = -1.0 / piv;
        if (li > k) a[k] = a[k] * t;
        // row elimination with column indexing
#pragma unroll
        for (int j = k + 1; j < NS; ++j) {
            const double tj = shfl_d(mask, a[j], l, G);
            const double akj = shfl_d(mask, a[j], k, G);
            if (l != k) {
                if (li == l) a[j] = akj;
                if (li == k) a[j] = tj;
            }
            if (li > k) a[j] = a[j] + tj * a[k];
        }
    }
    const double ann = shfl_d(mask, a[NS - 1], NS - 1, G);
    if (ann == 0.0) info = NS;
    return info;
}

// ---- dgefa with logical row interchanges ----
// Same factorisation, same arithmetic, but the data of a row never leaves its lane: LINPACK's
// interchange of rows l and k (columns >= k only, src/linpack.f90:404-420) becomes an update of
// the lanes' logical row numbers `r`.  That removes the two exchange shuffles per eliminated
// element; what is left is one broadcast of the pivot row's element per column.  The pivot search
// (idamax: first maximum) runs on the integer image of |a| with three warp reductions instead of a
// five-round shuffle butterfly.  Column k is final once step k is done; `store(col, pos, value)`
// is called for every lane with pos = the row LINPACK would hold that value in.
template <int NS, bool SPEC = true, class Store>
__device__ __forceinline__ int group_dgefa_track(double (&a)[NS], int li, int &ipv, Store store)
{
    constexpr int G = GroupOf<NS>::G;
    const unsigned mask = group_mask<G>();
    const unsigned lane0 = (threadIdx.x & 31u) & ~(unsigned)(G - 1);   // first warp lane of this group
    int info = 0;
    int r = li;   // logical (LINPACK) row currently held by this lane, for columns >= k
    ipv = NS;
#pragma unroll
    for (int k = 0; k < NS - 1; ++k) {
        const bool cand = (li < NS) && (r >= k);
        // every lane forms -1/a(i,k) for its own row while the pivot search runs; the pivot row's
        // value (== -1/a(k,k) after the interchange, linpack.f90:409) is broadcast afterwards, which
        // takes the divide latency off the critical path of the step
        const double tcand = SPEC ? -1.0 / a[k] : 0.0;
        const unsigned long long key = cand ? (unsigned long long)__double_as_longlong(fabs(a[k])) : 0ull;
        const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
        const unsigned mhi = __reduce_max_sync(mask, hi);
        const unsigned mlo = __reduce_max_sync(mask, (cand && hi == mhi) ? lo : 0u);
        const unsigned l = __reduce_min_sync(mask, (cand && hi == mhi && lo == mlo) ? (unsigned)r : 0xffu);
        const unsigned who = __ballot_sync(mask, cand && r == (int)l);
        const int pl = (int)(__ffs(who) - 1 - lane0);   // group-relative lane of the pivot row
        if (li == k) ipv = (int)l + 1;
        const double piv = shfl_d(mask, a[k], pl, G);
        if (piv == 0.0) {   // zero pivot: no interchange, column left as is (linpack.f90:394-397)
            info = k + 1;
            store(k, r, a[k]);
            continue;
        }
        if (r == (int)l) r = k;
        else if (r == k) r = (int)l;
        const double t = SPEC ? shfl_d(mask, tcand, pl, G) : -1.0 / piv;
        const bool below = r > k;           // rows still to be eliminated in this step
        if (below) a[k] = a[k] * t;         // multipliers (negated)
        store(k, r, a[k]);
        // broadcast the pivot row's tail first (all lanes), then one divergent region of straight-line
        // multiply-adds for the rows below the pivot: no per-element selects
        double tj[NS];
#pragma unroll
        for (int j = k + 1; j < NS; ++j) tj[j] = shfl_d(mask, a[j], pl, G);
        if (below) {
#pragma unroll
            for (int j = k + 1; j < NS; ++j) a[j] = a[j] + tj[j] * a[k];
        }
    }
    store(NS - 1, r, a[NS - 1]);
    const unsigned wl = __ballot_sync(mask, (li < NS) && r == NS - 1);
    const double ann = shfl_d(mask, a[NS - 1], (int)(__ffs(wl) - 1 - lane0), G);
    if (ann == 0.0) info = NS;
    return info;
}

// ---- dgesl (job = 0) for one block per group; b = this lane's b(i+1) ----
template <int NS>
__device__ __forceinline__ double group_dgesl(const double (&a)[NS], int ipv, int li, double b)
{
    constexpr int G = GroupOf<NS>::G;
    const unsigned mask = group_mask<G>();
    // forward: t = b(l); swap; b(k+1:n) += t*a(k+1:n,k)
#pragma unroll
    for (int k = 0; k < NS - 1; ++k) {
        const int l = __shfl_sync(mask, ipv, k, G) - 1;
        const double t = shfl_d(mask, b, l, G);
        const double bk = shfl_d(mask, b, k, G);
        if (l != k) {
            if (li == l) b = bk;
            if (li == k) b = t;
        }
        if (li > k && li < NS) b = b + t * a[k];
    }
    // back: b(k) /= a(k,k); b(1:k-1) += -b(k)*a(1:k-1,k)
#pragma unroll
    for (int k = NS - 1; k >= 0; --k) {
        if (li == k) b = b / a[k];
        const double t = -shfl_d(mask, b, k, G);
        if (li < k) b = b + t * a[k];
    }
    return b;
}

template <int NS>
__device__ __forceinline__ void group_load_block(const double *__restrict__ blk, int li, double (&a)[NS])
{
#pragma unroll
    for (int j = 0; j < NS; ++j) a[j] = (li < NS) ? __ldg(blk + j * NS + li) : 0.0;
}

