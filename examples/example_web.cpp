// examples/example_web.cpp -- the reference's food-web demonstration program (example/example_web.f90:721-990,
// original/examples/dweb.f) as a compiled host program on top of the C ABI of libdaskr_b200.so.
//
// What stays as in the reference: the DASKR calling sequence (info / rwork / iwork, IC call, loop over output times,
// root handling with idid = 5), the preconditioner choice jpre = ipar(1) = 3 (A_S * A_R), the two Krylov methods
// METH = 1 (no block grouping, setup_rbdpre) and METH = 2 (5 x 5 groups, setup_rbgpre), the table that is printed.
// What changes: res / jac / psol / rt are the library's device routines, c / cdot are host arrays that DASKR returns at
// every output time, rwork / iwork hold the scalar headers only.  METH = 0 (direct banded solver) is not part of this
// library.  Mesh and species count are run-time arguments (the reference fixes them as parameters at :11).
//
//   g++ -O2 -Iinclude examples/example_web.cpp -Ldaskr_b200/_lib -ldaskr_b200 -Wl,-rpath,$PWD/daskr_b200/_lib -o example_web
//   ./example_web [mx my np]          (defaults 20 20 1: the original demonstration problem)
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "daskr_b200.h"

int main(int argc, char **argv)
{
    const int mx = argc > 1 ? atoi(argv[1]) : 20, my = argc > 2 ? atoi(argv[2]) : 20, np = argc > 3 ? atoi(argv[3]) : 1;
    const int ns = 2 * np, neq = ns * mx * my, nrt = 1, nout = 18;
    if (dkb_init(0) != 0) {
        fprintf(stderr, "dkb_init: %s\n", dkb_last_error());
        return 1;
    }
    printf("DWEB: Example program for DASKR package (device path)\n\n");
    printf("Food web problem with NS species, NS =%4d\nPredator-prey interaction and diffusion on a 2-D square\n\n", ns);
    printf("Mesh dimensions (MX,MY) =%5d%5d\nTotal system size is NEQ =%8d\n\n", mx, my, neq);
    printf("Root function is R(Y) = average(c1) - 20\n\n");

    std::vector<double> c(neq), cdot(neq), rwork(60 + 3 * nrt), rpar(1);
    std::vector<int> iwork(40 + neq);
    const int64_t lrw = (int64_t)rwork.size(), liw = (int64_t)iwork.size();
    double rtol = 1e-5, atol = 1e-5;
    int info[20], ipar[2], jroot[1], idid = 0, status = 0;

    for (int meth = 1; meth <= 2; ++meth) {
        for (int i = 0; i < 20; ++i) info[i] = 0;
        info[10] = 1;   // info(11): compute consistent initial values
        info[11] = 1;   // info(12): Krylov
        info[13] = 1;   // info(14): stop after the IC calculation
        info[14] = 1;   // info(15): jac routine supplied
        info[15] = 1;   // info(16): error test on the differential variables only
        const int jbg = meth - 1, jpre = 3, lid = 40;
        ipar[0] = jpre;
        ipar[1] = jbg;
        printf("\n%s\n\nLinear solver method flag INFO(12) = 1  (0 = direct, 1 = Krylov)\n", std::string(80, '.').c_str());
        printf("Preconditioner flag is JPRE =%3d\n(1 = reaction factor A_R, 2 = spatial factor A_S, 3 = A_S*A_R, 4 = A_R*A_S)\n", jpre);
        for (auto &v : iwork) v = 0;
        if (jbg == 0) {
            dkb_setup_rbdpre(&mx, &my, &ns, &np, &lid, iwork.data());
            printf("No block-grouping in reaction factor\n");
        } else {
            const int nxg = 5, nyg = 5;
            dkb_setup_rbgpre(&mx, &my, &ns, &np, &nxg, &nyg, &lid, iwork.data());
            printf("Block-grouping in reaction factor\nNumber of groups =%5d (NGX by NGY, NGX =%3d,  NGY =%3d)\n", nxg * nyg, nxg, nyg);
        }
        if (dkb_web_setpar(np, mx, my) != 0 || dkb_web_cinit(c.data(), cdot.data(), 1e5 * np) != 0) {
            fprintf(stderr, "%s\n", dkb_last_error());
            return 1;
        }
        double t = 0.0, tout = 1e-8;
        int nli = 0, nni = 0;
        printf("\n          t      <c1>  NSTEP   NRE   NNI   NLI   NPE    NQ          H    AVLIN\n");
        for (int iout = 0; iout <= nout; ++iout) {
            for (;;) {
                dkb_daskr(dkb_web_res, &neq, &t, c.data(), cdot.data(), &tout, info, &rtol, &atol, &idid, rwork.data(), &lrw,
                          iwork.data(), &liw, rpar.data(), ipar, dkb_web_jac, dkb_web_psol, (void *)dkb_web_rt, &nrt, jroot);
                const int nst = iwork[10], nre = iwork[11], npe = iwork[12], nqu = iwork[7];
                const int nnidif = iwork[18] - nni, nlidif = iwork[19] - nli;
                nni = iwork[18];
                nli = iwork[19];
                const double hu = rwork[6], avlin = nnidif > 0 ? (double)nlidif / nnidif : 0.0;
                double c1ave = 0.0;   // c1_average, example_web.f90:577-598
                for (int p = 0; p < mx * my; ++p) c1ave += c[(size_t)p * ns];
                c1ave /= (double)mx * my;
                printf("%11.5e%10.5f%7d%6d%6d%6d%6d%6d%11.2e%9.4f\n", t, c1ave, nst, nre, nni, nli, npe, nqu, hu, avlin);
                if (idid == 5) printf("               *****   Root found, JROOT =%3d\n", jroot[0]);
                else break;
            }
            if (idid < 0) {
                printf("\n\nFinal time reached =%12.4e   (idid = %d: %s)\n\n", t, idid, dkb_last_error());
                status = 2;
                break;
            }
            tout = (tout > 0.9) ? tout + 1.0 : 10 * tout;
            if (iout == 0) {
                info[10] = 0;
                tout = 1e-8;
                nli = 0;
                nni = 0;
            }
        }
        const int nst = iwork[10], nre = iwork[11], npe = iwork[12], nps = iwork[20];
        const double avdim = nni > 0 ? (double)iwork[19] / iwork[18] : 0.0;
        printf("\nFinal statistics for this run..\n");
        printf("Number of time steps              =%6d\nNumber of residual evaluations    =%6d\n", nst, nre);
        printf("Number of root fn. evaluations    =%6d\nNumber of precond. evaluations    =%6d\n", iwork[35], npe);
        printf("Number of precond. solves         =%6d\nNumber of nonlinear iterations    =%6d\n", nps, iwork[18]);
        printf("Number of linear iterations       =%6d\nAverage Krylov subspace dimension =%8.4f\n", iwork[19], avdim);
        printf("%d nonlinear conv. failures,%5d linear conv. failures\n", iwork[14], iwork[15]);
    }
    dkb_finalize();
    return status;
}
