/*
 * oracle/blas_linpack.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Level-1 BLAS as vendored by the reference in extern/dblas.f (the reference
 * links OpenBLAS when present, fpm.toml:13; the vendored file is the
 * arithmetic source of truth) and LINPACK dgefa/dgesl from
 * src/linpack.f90.  Stride-1 forms only: that is all the Krylov path uses.
 */
#include "oracle.h"
#include <cmath>

extern "C" {

/* extern/dblas.f:138-185.  The 5-way unrolled body accumulates left to right
 * into the single accumulator dtemp, so it is the in-order sum. */
double o_ddot(long n, const double *dx, const double *dy)
{
    double dtemp = 0.0;
    if (n <= 0) return 0.0;
    long m = n % 5;
    for (long i = 0; i < m; ++i) dtemp = dtemp + dx[i] * dy[i];
    if (n < 5) return dtemp;
    for (long i = m; i < n; i += 5)
        dtemp = dtemp + dx[i] * dy[i] + dx[i + 1] * dy[i + 1] + dx[i + 2] * dy[i + 2] +
                dx[i + 3] * dy[i + 3] + dx[i + 4] * dy[i + 4];
    return dtemp;
}

/* extern/dblas.f:186-307 -- Lawson's four-phase scaled 2-norm.  The assigned
 * GOTO of the original is restated as the phase variable `next`. */
double o_dnrm2(long n, const double *dx)
{
    const double cutlo = 8.232e-11, cuthi = 1.304e19;
    if (n <= 0) return 0.0;
    int next = 30;
    double sum = 0.0, xmax = 0.0;
    long i = 0;
    for (;;) {
        double xi = dx[i];
        bool to85 = false;
        switch (next) {
        case 30:
            if (std::fabs(xi) > cutlo) { to85 = true; break; }
            next = 50;
            xmax = 0.0;
            /* fall through */
        case 50:
            if (xi == 0.0) goto L200;
            if (std::fabs(xi) > cutlo) { to85 = true; break; }
            next = 70;
            xmax = std::fabs(xi); /* 105 */
            sum = sum + (xi / xmax) * (xi / xmax); /* 115 */
            goto L200;
        case 70:
            if (std::fabs(xi) > cutlo) { /* 75 */
                sum = (sum * xmax) * xmax;
                to85 = true;
                break;
            }
            /* fall through */
        case 110:
            if (std::fabs(xi) <= xmax) {
                sum = sum + (xi / xmax) * (xi / xmax); /* 115 */
            } else {
                sum = 1.0 + sum * ((xmax / xi) * (xmax / xi));
                xmax = std::fabs(xi);
            }
            goto L200;
        }
        if (to85) {
            /* 85: mid-range phase */
            double hitest = cuthi / (double)(float)n;
            long j = i;
            for (; j < n; ++j) {
                if (std::fabs(dx[j]) >= hitest) break;
                sum = sum + dx[j] * dx[j];
            }
            if (j >= n) return std::sqrt(sum);
            /* 100: switch to the large-value phase */
            i = j;
            next = 110;
            sum = (sum / dx[i]) / dx[i];
            xmax = std::fabs(dx[i]); /* 105 */
            sum = sum + (dx[i] / xmax) * (dx[i] / xmax); /* 115 */
        }
    L200:
        ++i;
        if (i >= n) break;
    }
    return xmax * std::sqrt(sum);
}

/* extern/dblas.f:1-47 (note the early return for da == 0) */
void o_daxpy(long n, double da, const double *dx, double *dy)
{
    if (n <= 0) return;
    if (da == 0.0) return;
    for (long i = 0; i < n; ++i) dy[i] = dy[i] + da * dx[i];
}

/* extern/dblas.f:97-137 */
void o_dscal(long n, double da, double *dx)
{
    for (long i = 0; i < n; ++i) dx[i] = da * dx[i];
}

/* extern/dblas.f:48-96 */
void o_dcopy(long n, const double *sx, double *sy)
{
    for (long i = 0; i < n; ++i) sy[i] = sx[i];
}

/* extern/dblas.f:308-344 -- strict '>' so ties resolve to the lowest index;
 * returns a 1-based index like the Fortran function. */
int o_idamax(int n, const double *dx)
{
    if (n < 1) return 0;
    int imax = 1;
    if (n == 1) return imax;
    double dmax = std::fabs(dx[0]);
    for (int i = 2; i <= n; ++i) {
        if (std::fabs(dx[i - 1]) <= dmax) continue;
        imax = i;
        dmax = std::fabs(dx[i - 1]);
    }
    return imax;
}

#define A(i, j) a[((long)(j) - 1) * lda + ((i) - 1)]

/* src/linpack.f90:325-431 */
void o_dgefa(double *a, int lda, int n, int *ipvt, int *info)
{
    *info = 0;
    for (int k = 1; k <= n - 1; ++k) {
        int l = o_idamax(n - k + 1, &A(k, k)) + k - 1;
        ipvt[k - 1] = l;
        if (A(l, k) == 0.0) { /* zero pivot: column already triangularised */
            *info = k;
            continue;
        }
        if (l != k) {
            double t = A(l, k);
            A(l, k) = A(k, k);
            A(k, k) = t;
        }
        double t = -1.0 / A(k, k);
        for (int i = k + 1; i <= n; ++i) A(i, k) = A(i, k) * t;
        for (int j = k + 1; j <= n; ++j) {
            t = A(l, j);
            if (l != k) {
                A(l, j) = A(k, j);
                A(k, j) = t;
            }
            for (int i = k + 1; i <= n; ++i) A(i, j) = A(i, j) + t * A(i, k);
        }
    }
    ipvt[n - 1] = n;
    if (A(n, n) == 0.0) *info = n;
}

/* src/linpack.f90:433-546 */
void o_dgesl(const double *a, int lda, int n, const int *ipvt, double *b, int job)
{
    if (job == 0) {
        for (int k = 1; k <= n - 1; ++k) {
            int l = ipvt[k - 1];
            double t = b[l - 1];
            if (l != k) {
                b[l - 1] = b[k - 1];
                b[k - 1] = t;
            }
            for (int i = k + 1; i <= n; ++i) b[i - 1] = b[i - 1] + t * A(i, k);
        }
        for (int k = n; k >= 1; --k) {
            b[k - 1] = b[k - 1] / A(k, k);
            double t = -b[k - 1];
            for (int i = 1; i <= k - 1; ++i) b[i - 1] = b[i - 1] + t * A(i, k);
        }
    } else {
        for (int k = 1; k <= n; ++k) {
            double t = 0.0;
            for (int i = 1; i <= k - 1; ++i) t = t + A(i, k) * b[i - 1];
            b[k - 1] = (b[k - 1] - t) / A(k, k);
        }
        for (int k = n - 1; k >= 1; --k) {
            double t = 0.0;
            for (int i = k + 1; i <= n; ++i) t = t + A(i, k) * b[i - 1];
            b[k - 1] = b[k - 1] + t;
            int l = ipvt[k - 1];
            if (l != k) {
                t = b[l - 1];
                b[l - 1] = b[k - 1];
                b[k - 1] = t;
            }
        }
    }
}
#undef A

/* Blocks laid out as jac_rbdpre leaves them (src/daskr_rbdpre.f90:237-249):
 * block p at bd + p*ns*ns column-major, pivots at ipbd + p*ns, 1-based.
 * info[p] is dgefa's info for block p (jac_rbdpre itself stops at the first
 * non-zero one; the micro-sweep of BASELINE config 5 factors them all). */
void o_dgefa_batched(double *bd, int *ipbd, int ns, long nblk, int *info)
{
    for (long p = 0; p < nblk; ++p)
        o_dgefa(bd + p * (long)ns * ns, ns, ns, ipbd + p * ns, info + p);
}

/* src/daskr_rbdpre.f90:269-277 */
void o_dgesl_batched(const double *bd, const int *ipbd, int ns, long nblk, double *b)
{
    for (long p = 0; p < nblk; ++p)
        o_dgesl(bd + p * (long)ns * ns, ns, ns, ipbd + p * ns, b + p * ns, 0);
}

} /* extern "C" */
