// rbdpre_user.cu -- jac_rbdpre (src/daskr_rbdpre.f90:158-251) for ANY user reaction term.
//
// The reference calls the user's `rblock(t, jx, jy, uxy, rxy)` once per mesh point and perturbed species
// (src/daskr_rbdpre.f90:176-189, 219-229).  On the device that hook is batched: the user routine evaluates R for all
// mesh points of a state in one call (one kernel launch on dkb_stream()),
//
//     rblock(t, mx, my, ns, u, r, user)       u, r: device arrays of ns*mx*my doubles, species fastest
//
// and jac_rbdpre becomes ns rounds of { perturb species js of every point, rblock, difference quotient column js of
// every block, restore }, followed by "+ cj on the first nsd diagonals" and the batched dgefa -- the same arithmetic per
// point as the reference (del = max(srur*|u_j|, 0.01/rewt_j), column = (r1 - r0)*(-1/del)), so blocks and pivots are
// bit-identical to the CPU path for a bit-identical R.  Output is the reference's own bd / ipbd layout (block p at
// bd + p*ns*ns, column-major; pivots 1-based at ipbd + p*ns), which dkb_psol_rbdpre solves with directly;
// dkb_rbd_from_reference converts it to the tile-interleaved layout of the streaming solve (dkb_psol_rbdpre_tiles).
// The food-web and heat problems do not come through here: their blocks are built by one fused kernel (jac2.cu, heat.cu).
#include "dkb_internal.h"
#include <limits.h>

// round js, before rblock: save u_js, form del, perturb in place (:219-224)
__global__ void __launch_bounds__(256) rbd_perturb_kernel(i64 npts, int ns, int js, double srur, double *__restrict__ u,
                                                          const double *__restrict__ rewt, double *__restrict__ usave,
                                                          double *__restrict__ delv)
{
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += stride) {
        const i64 j = p * ns + js;
        const double uj = u[j];
        const double del = fmax(srur * fabs(uj), 1e-2 / rewt[j]);
        usave[p] = uj;
        delv[p] = del;
        u[j] = uj + del;
    }
}

// round js, after rblock: column js of every block (:225-228), restore u_js (:229)
__global__ void __launch_bounds__(256) rbd_column_kernel(i64 npts, int ns, int js, const double *__restrict__ r1,
                                                         const double *__restrict__ r0, const double *__restrict__ delv,
                                                         const double *__restrict__ usave, double *__restrict__ u,
                                                         double *__restrict__ bd)
{
    const i64 n = npts * ns;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += stride) {
        const i64 p = g / ns;
        const int i = (int)(g - p * ns);
        const double fac = -1.0 / delv[p];
        bd[p * ns * ns + (i64)js * ns + i] = (r1[g] - r0[g]) * fac;
        if (i == js) u[g] = usave[p];
    }
}

// :240-244: cj on the first nsd diagonal entries of every block
__global__ void __launch_bounds__(256) rbd_add_cj_kernel(i64 npts, int ns, int nsd, double cj, double *__restrict__ bd)
{
    const i64 n = npts * nsd;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 g = (i64)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += stride) {
        const i64 p = g / nsd;
        const int i = (int)(g - p * nsd);
        double *d = bd + p * ns * ns + (i64)i * (ns + 1);
        *d = *d + cj;
    }
}

extern "C" {

// jac_rbdpre(t, u, r0, rblock, r1, rewt, cj, bd, ipbd, ierr), src/daskr_rbdpre.f90:158-251, with the batched rblock hook
// (+ its `user` pointer).  u is perturbed in place and restored, like the reference; r1 is N doubles of scratch (the
// reference needs ns: one point at a time).  *ierr = dgefa's info of the first failing block, 0 if none -- the reference
// stops factoring there, here the remaining blocks are factored as well.
void dkb_jac_rbdpre(const double *t, double *u, const double *r0, dkb_rblock_t rblock, void *user, double *r1,
                    const double *rewt, const double *cj, double *bd, int *ipbd, int *ierr)
{
    DkbCtx *c = g_ctx;
    *ierr = -1;
    if (!c || !c->rbd_set || rblock == nullptr) {
        dkb_fail(DKB_E_STATE, "dkb_jac_rbdpre: needs dkb_init, dkb_setup_rbdpre and an rblock routine");
        return;
    }
    const int ns = c->ns, nsd = c->nsd;
    const i64 npts = c->npts;
    double *usave = nullptr, *delv = nullptr;
    int *info = nullptr;
    if (cudaMalloc(&usave, sizeof(double) * npts) != cudaSuccess || cudaMalloc(&delv, sizeof(double) * npts) != cudaSuccess ||
        cudaMalloc(&info, sizeof(int) * npts) != cudaSuccess) {
        cudaFree(usave);
        cudaFree(delv);
        dkb_fail(DKB_E_CUDA, "dkb_jac_rbdpre: cannot allocate %lld bytes of scratch", (long long)(20 * npts));
        return;
    }
    const double srur = 1.4901161193847656e-08;   // sqrt(epsilon(one)) = 2^-26, :129
    const int mx = c->mx, my = c->my;
    const int gp = (int)std::min<i64>((npts + 255) / 256, (i64)c->nsm * 8);
    const int gn = (int)std::min<i64>((npts * ns + 255) / 256, (i64)c->nsm * 8);
    for (int js = 0; js < ns; ++js) {
        rbd_perturb_kernel<<<gp, 256, 0, c->stream>>>(npts, ns, js, srur, u, rewt, usave, delv);
        rblock(t, &mx, &my, &ns, u, r1, user);
        rbd_column_kernel<<<gn, 256, 0, c->stream>>>(npts, ns, js, r1, r0, delv, usave, u, bd);
        c->launches += 2;
    }
    if (nsd > 0) {
        const int gd = (int)std::min<i64>((npts * nsd + 255) / 256, (i64)c->nsm * 8);
        rbd_add_cj_kernel<<<gd, 256, 0, c->stream>>>(npts, ns, nsd, *cj, bd);
        c->launches += 1;
    }
    c->h_iflag[2] = INT_MAX;
    cudaMemcpyAsync(c->d_iflag + 2, c->h_iflag + 2, sizeof(int), cudaMemcpyHostToDevice, c->stream);
    int rc = lu_dgefa_batched(bd, ipbd, ns, npts, info, c->d_iflag + 2);
    allreduce_min_int(c->d_iflag + 2, 1);
    fetch_words(c->d_iflag + 2, c->h_iflag + 2, sizeof(int));
    if (rc == 0 && cudaGetLastError() == cudaSuccess) {
        *ierr = 0;
        if (c->h_iflag[2] != INT_MAX && c->h_iflag[2] >= 1 && c->h_iflag[2] <= npts) {   // first failing block (this rank's)
            int k = 1;
            cudaMemcpy(&k, info + (c->h_iflag[2] - 1), sizeof(int), cudaMemcpyDeviceToHost);
            *ierr = k;
        } else if (c->h_iflag[2] != INT_MAX) {
            *ierr = 1;   // a block on another rank failed
        }
    }
    cudaFree(usave);
    cudaFree(delv);
    cudaFree(info);
}

// psol_rbdpre (src/daskr_rbdpre.f90:253-279) on the tile-interleaved copy of the blocks (dkb_rbd_from_reference, or what
// the built-in jac routines write): the streaming thread-per-point solve of psol_tpp.cu
void dkb_psol_rbdpre_tiles(double *b, const double *rwp, const int *iwp)
{
    DkbCtx *c = g_ctx;
    if (!c || !c->rbd_set) {
        fprintf(stderr, "daskr_b200: dkb_psol_rbdpre_tiles before dkb_setup_rbdpre\n");
        abort();
    }
    if (c->ns == 1) lu_dgesl_batched(rwp, iwp, 1, c->npts, b);
    else if (tpp_psol(rwp, iwp, c->ns, c->npts, b, nullptr, nullptr, 1.0, b, -1) != 0) {
        fprintf(stderr, "daskr_b200: dkb_psol_rbdpre_tiles: %s\n", c->err);
        abort();
    }
}

} // extern "C"
