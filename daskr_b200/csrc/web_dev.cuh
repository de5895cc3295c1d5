// web_dev.cuh -- device-side view of the food-web problem parameters (shared by web.cu and
// datv_fused.cu).  Built by make_webdev() from the context (module web_par, example_web.f90:5-60).
#pragma once
#include "dkb_internal.h"

struct WebDev {
    int ns, np, mx, my, jy0, my_g, has_lo, has_hi;
    double alpha, beta, dx, dy, srur;
    const double *acoef, *bcoef, *cox, *coy, *sx, *sy;   // device tables
};

WebDev make_webdev();
// copies acoef / bcoef / cox / coy into the constant bank the fused datv kernel reads them from
void datv_fused_upload_tables(const WebPar &p);
void jac2_upload_tables(const WebPar &p);
// residual only, same staging as the single datv kernel (datv_fused.cu): mode 1 = perturbed - savr, 2 = plain res
bool web_resid_single(const WebDev &w, int mode, const double *v, double vscale, const double *wt, double wscale, const double *y,
                      const double *ydot, const double *savr, double cj, double *out);
// two-phase FD + LU block Jacobian (jac2.cu)
bool web_jac2(const WebDev &w, int nsd, const double *c, const double *rewt, double cj, double *bd, int *ipbd,
              int *d_first_flag);
// same result, two rows of a point's block per lane for both phases (jac3.cu); false if ns is not instantiated
bool web_jac3(const WebDev &w, int nsd, const double *c, const double *rewt, double cj, double *bd, int *ipbd,
              int *d_first_flag);
// single-kernel datv; returns false if the mesh / ns is not eligible (caller falls back)
bool web_datv_single(const WebDev &w, const double *v, double vscale, const double *wt, double wscale, const double *y,
                     const double *ydot, const double *savr, double cj, const double *lut, const int *pt,
                     double *vnext);
