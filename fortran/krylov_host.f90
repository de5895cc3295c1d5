!> krylov_host.f90 -- how DASKR's own Fortran host logic calls the device path (uncompiled: no Fortran front end here).
!>
!> BASELINE's north star keeps the BDF / Newton control of the reference (ddstp, dnedk, dnsk, DDASIC) in Fortran and
!> moves every statement on a length-NEQ array behind the C ABI.  This file shows that patch on the Newton iteration
!> (dnsk, src/daskr_new.f90:2881-3034) and on the linear solver entry (dslvk, :3036-3165); the remaining routines
!> change the same way, statement by statement, with the table at the end.  The vectors are device addresses
!> (type(c_ptr)) obtained once from dkb_workspace_get; scalars, counters and all control flow stay what they are.
module daskr_device_host
   use, intrinsic :: iso_c_binding
   use daskr_b200
   implicit none
   integer, parameter :: rk = c_double
   ! iwm slots, src/daskr_new.f90:100-145
   integer, parameter :: iwloc_nre = 12, iwloc_nni = 19

contains

   !> dslvk with the reference's argument list (rwm is gone: the Krylov basis lives in the device workspace)
   subroutine dslvk(neq, y, t, ydot, savr, x, ewt, iwm, res, ires, psol, iersl, cj, epslin, sqrtn, rsqrtn, rhok, rpar, ipar)
      integer(c_int), intent(in) :: neq
      type(c_ptr), intent(in) :: y, ydot, savr, x, ewt
      real(rk), intent(in) :: t, cj, epslin, sqrtn, rsqrtn
      integer(c_int), intent(inout) :: iwm(*), ires, iersl, ipar(*)
      type(c_funptr), intent(in) :: res, psol
      real(rk), intent(inout) :: rhok, rpar(*)
      call dslvk_dev(neq, y, t, ydot, savr, x, ewt, iwm, res, ires, psol, iersl, cj, epslin, sqrtn, rsqrtn, rhok, rpar, ipar)
   end subroutine dslvk

   !> dnsk: the Newton iteration of one BDF step.  Same arguments as the reference minus rwm; res / psol are c_funptr.
   subroutine dnsk(t, y, ydot, neq, res, psol, wt, rpar, ipar, savr, delta, e, iwm, cj, sqrtn, rsqrtn, epslin, epscon, &
                   s, confac, tolnew, muldel, maxit, ires, iersl, iernew)
      real(rk), intent(in) :: t, cj, sqrtn, rsqrtn, epslin, epscon, confac, tolnew
      type(c_ptr), intent(in) :: y, ydot, wt, savr, delta, e
      integer(c_int), intent(in) :: neq, muldel, maxit
      type(c_funptr), intent(in) :: res, psol
      real(rk), intent(inout) :: rpar(*), s
      integer(c_int), intent(inout) :: ipar(*), iwm(*)
      integer(c_int), intent(out) :: ires, iersl, iernew
      procedure(dkb_res_t), pointer :: res_f
      integer(c_int64_t) :: n
      integer(c_int) :: rc, m
      real(rk) :: deltanorm, oldnorm, rate, rhok
      logical :: converged

      n = int(neq, c_int64_t)
      call c_f_procpointer(res, res_f)
      m = 0
      oldnorm = 0
      converged = .false.
      rc = dkb_vscale(n, 0.0_rk, e)                               ! e = zero                      (:2966)
      newton: do
         iwm(iwloc_nni) = iwm(iwloc_nni) + 1
         if (muldel == 1) rc = dkb_vscale(n, confac, delta)       ! delta = delta*confac          (:2975)
         rc = dkb_vcopy(n, delta, savr)                           ! savr = delta                  (:2979)
         call dslvk(neq, y, t, ydot, savr, delta, wt, iwm, res, ires, psol, iersl, cj, epslin, sqrtn, rsqrtn, rhok, &
                    rpar, ipar)                                   !                               (:2982)
         if (ires /= 0 .or. iersl /= 0) exit newton
         rc = dkb_newton_update(n, delta, cj, y, e, ydot)         ! y, e, ydot updates            (:2988-2990)
         deltanorm = ddwnrm_dev(neq, delta, wt)                   !                               (:2993)
         if (m == 0) then
            oldnorm = deltanorm
            converged = deltanorm <= tolnew
            if (converged) exit newton
         else
            rate = (deltanorm/oldnorm)**(1.0_rk/m)
            if (rate > 0.9_rk) exit newton
            s = rate/(1.0_rk - rate)
         end if
         converged = s*deltanorm <= epscon
         if (converged) exit newton
         m = m + 1
         if (m >= maxit) exit newton
         ires = 0
         call res_f(t, y, ydot, cj, delta, ires, rpar, ipar)      ! device res                    (:3018)
         iwm(iwloc_nre) = iwm(iwloc_nre) + 1
         if (ires < 0) exit newton
      end do newton
      if (converged) then
         iernew = 0
      else if (ires <= -2 .or. iersl < 0) then
         iernew = -1
      else
         iernew = 1
      end if
   end subroutine dnsk

end module daskr_device_host

! Statement table for the other routines (reference line -> call; n = neq, arrays are device addresses):
!   daskr.f:1750-1751, 1860-1861, 1876-1877, 2099-2101   ddawts + dinvwt         dkb_ewt_set_invert(n, info(2), rtol, atol, y, wt, vt, id, ier)
!   daskr.f:1755, 1881, 2106-2109                        vt = max(id,0)*wt       dkb_mask_vt(n, id, wt, vt)   (or vt in the call above)
!   daskr.f:1800, 1894, 2113; ddstp:360, 368, 376, 409   ddwnrm                  ddwnrm_dev(neq, v, wt)  /  dkb_error_norms (the three of ddstp in one pass)
!   daskr.f:1918-1920                                    phi(:,1) = y, phi(:,2) = h*yprime      dkb_vcopy, dkb_vscale_to
!   daskr.f:1975 ... 2237, DRCHEK                         ddatrp                  dkb_interp(n, t, tout, kold, phi, ldphi, psi, y, yprime)
!   ddstp:335                                            phi(:,j) = beta(j)*phi(:,j)            dkb_vscale(n, beta(j), phi_j)
!   ddstp:459-466                                        phi update cascade      dkb_phi_update(n, kp1, write_kp2, phi, ldphi, e)
!   ddstp:482, 552, 562                                  phi(:,j) = phi(:,j)/beta(j) ...        dkb_vscale / dkb_lincomb
!   dnedk:2808-2813                                      predictor               dkb_predict(n, kp1, phi, ldphi, gama, y, yprime)
!   dnedk:2828                                           call jac(...)           user jac on device arrays: dkb_web_jac, or dkb_jac_rbdpre + a batched rblock
!   dnedk:2854-2857                                      delta = min(y,0); e = e - delta        dkb_lincomb (after the user's projection kernel)
!   dnsk:2966-2993                                       see dnsk above
!   dslvk:3120-3163, dspigm, datv, dorth                 the whole GMRES solve    dslvk_dev (dkb_dslvk)
!   dnsik:2275-2338, dlinsk:2489-2546, dfnrmk:2646-2656  IC path                 dkb_vcopy / dkb_lincomb / ddwnrm_dev + user res / psol
