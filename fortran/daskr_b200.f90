!> daskr_b200.f90 -- ISO_C_BINDING interfaces of libdaskr_b200.so (include/daskr_b200.h) for a Fortran host.
!>
!> The reference (HugoMVale/daskr) has no FFI layer: its boundary is the Fortran-77 procedure convention.  The C ABI
!> keeps that convention (every argument by reference, INTEGER = c_int, reals = c_double, callbacks = c_funptr), so
!> each interface below is the reference routine's own argument list with length-NEQ arrays turned into device
!> addresses (type(c_ptr), by value).  Two ways to use it:
!>   1. replace DASKR as a whole:  call daskr(...) -> dkb_daskr, with setup_rbdpre / the built-in food-web routines
!>      (INTEGRATION.md section 2 shows the three changed lines of example_web.f90);
!>   2. keep DASKR's own driver (ddstp / dnedk / dnsk in Fortran, as BASELINE's north star words it) and route its
!>      array statements through the vector-op layer and dslvk through dkb_dslvk: fortran/krylov_host.f90.
!> NOT COMPILED in this repository's image (no Fortran front end: gfortran / flang / nvfortran are all absent); the same
!> entry points are exercised through ctypes (tests/test_gpu_*.py) and from C++ (examples/example_web.cpp).
module daskr_b200
   use, intrinsic :: iso_c_binding
   implicit none
   private :: c_int, c_double, c_int64_t, c_ptr, c_funptr, c_char

   integer(c_int), parameter :: DKB_OPT_DEVICE_IO = 1, DKB_OPT_FUSED = 2, DKB_OPT_MAX_STEPS = 3, DKB_OPT_ASYNC_OUT = 4

   abstract interface
      !> res_t, src/daskr_new.f90:11-21 -- y, ydot, delta are device addresses
      subroutine dkb_res_t(t, y, ydot, cj, delta, ires, rpar, ipar) bind(C)
         import :: c_double, c_int, c_ptr
         real(c_double), intent(in) :: t, cj
         type(c_ptr), value :: y, ydot, delta
         integer(c_int), intent(inout) :: ires
         real(c_double), intent(inout) :: rpar(*)
         integer(c_int), intent(inout) :: ipar(*)
      end subroutine dkb_res_t
      !> jack_t, src/daskr_new.f90:35-53
      subroutine dkb_jack_t(res, ires, neq, t, y, ydot, rewt, savr, wk, h, cj, rwp, iwp, ierr, rpar, ipar) bind(C)
         import :: c_double, c_int, c_ptr, c_funptr
         type(c_funptr), value :: res
         integer(c_int), intent(inout) :: ires, ierr
         integer(c_int), intent(in) :: neq
         real(c_double), intent(in) :: t, h, cj
         type(c_ptr), value :: y, ydot, rewt, savr, wk, rwp, iwp
         real(c_double), intent(inout) :: rpar(*)
         integer(c_int), intent(inout) :: ipar(*)
      end subroutine dkb_jack_t
      !> psol_t, src/daskr_new.f90:55-71
      subroutine dkb_psol_t(neq, t, y, ydot, savr, wk, cj, wght, rwp, iwp, b, epslin, ierr, rpar, ipar) bind(C)
         import :: c_double, c_int, c_ptr
         integer(c_int), intent(in) :: neq
         real(c_double), intent(in) :: t, cj, epslin
         type(c_ptr), value :: y, ydot, savr, wk, wght, rwp, iwp, b
         integer(c_int), intent(inout) :: ierr
         real(c_double), intent(inout) :: rpar(*)
         integer(c_int), intent(inout) :: ipar(*)
      end subroutine dkb_psol_t
      !> rt, src/daskr.f:209-231 (root functions) -- y, yp are device addresses, rval(nrt) is a host array
      subroutine dkb_rt_t(neq, t, y, yp, nrt, rval, rpar, ipar) bind(C)
         import :: c_double, c_int, c_ptr
         integer(c_int), intent(in) :: neq, nrt
         real(c_double), intent(in) :: t
         type(c_ptr), value :: y, yp
         real(c_double), intent(out) :: rval(*)
         real(c_double), intent(inout) :: rpar(*)
         integer(c_int), intent(inout) :: ipar(*)
      end subroutine dkb_rt_t
      !> rblock of jac_rbdpre, src/daskr_rbdpre.f90:176-189, batched over the mesh: r(:) = R(t, u(:)) for all points
      subroutine dkb_rblock_batched_t(t, mx, my, ns, u, r, user) bind(C)
         import :: c_double, c_int, c_ptr
         real(c_double), intent(in) :: t
         integer(c_int), intent(in) :: mx, my, ns
         type(c_ptr), value :: u, r, user
      end subroutine dkb_rblock_batched_t
   end interface

   interface
      ! ---- context ----
      function dkb_init(device) bind(C, name="dkb_init") result(rc)
         import :: c_int
         integer(c_int), value :: device
         integer(c_int) :: rc
      end function
      function dkb_finalize() bind(C, name="dkb_finalize") result(rc)
         import :: c_int
         integer(c_int) :: rc
      end function
      function dkb_set_option(opt, val) bind(C, name="dkb_set_option") result(rc)
         import :: c_int, c_int64_t
         integer(c_int), value :: opt
         integer(c_int64_t), value :: val
         integer(c_int) :: rc
      end function
      function dkb_comm_unique_id(id128) bind(C, name="dkb_comm_unique_id") result(rc)
         import :: c_int, c_ptr
         type(c_ptr), value :: id128
         integer(c_int) :: rc
      end function
      function dkb_comm_init(id128, rank, nranks) bind(C, name="dkb_comm_init") result(rc)
         import :: c_int, c_ptr
         type(c_ptr), value :: id128
         integer(c_int), value :: rank, nranks
         integer(c_int) :: rc
      end function

      ! ---- setup_rbdpre / setup_rbgpre, src/daskr_rbdpre.f90:102-156, src/daskr_rbgpre.f90:107-158 ----
      subroutine setup_rbdpre(mx, my, ns, nsd, lid, iwork) bind(C, name="dkb_setup_rbdpre")
         import :: c_int
         integer(c_int), intent(in) :: mx, my, ns, nsd, lid
         integer(c_int), intent(inout) :: iwork(*)
      end subroutine
      subroutine setup_rbgpre(mx, my, ns, nsd, nxg, nyg, lid, iwork) bind(C, name="dkb_setup_rbgpre")
         import :: c_int
         integer(c_int), intent(in) :: mx, my, ns, nsd, nxg, nyg, lid
         integer(c_int), intent(inout) :: iwork(*)
      end subroutine
      function dkb_slab(jy0, myloc) bind(C, name="dkb_slab") result(rc)
         import :: c_int
         integer(c_int), intent(out) :: jy0, myloc
         integer(c_int) :: rc
      end function

      ! ---- jac_rbdpre / psol_rbdpre with a user reaction term, src/daskr_rbdpre.f90:158-279 ----
      subroutine jac_rbdpre(t, u, r0, rblock, user, r1, rewt, cj, bd, ipbd, ierr) bind(C, name="dkb_jac_rbdpre")
         import :: c_int, c_double, c_ptr, c_funptr
         real(c_double), intent(in) :: t, cj
         type(c_ptr), value :: u, r0, user, r1, rewt, bd, ipbd
         type(c_funptr), value :: rblock
         integer(c_int), intent(out) :: ierr
      end subroutine
      subroutine psol_rbdpre_tiles(b, bd, ipbd) bind(C, name="dkb_psol_rbdpre_tiles")
         import :: c_ptr
         type(c_ptr), value :: b, bd, ipbd
      end subroutine
      subroutine psol_rbdpre(b, bd, ipbd) bind(C, name="dkb_psol_rbdpre")   ! reference bd / ipbd layout
         import :: c_ptr
         type(c_ptr), value :: b, bd, ipbd
      end subroutine

      ! ---- DASKR, src/daskr.f:1-3 ----
      subroutine daskr(res, neq, t, y, yprime, tout, info, rtol, atol, idid, rwork, lrw, iwork, liw, &
                       rpar, ipar, jac, psol, rt, nrt, jroot) bind(C, name="dkb_daskr")
         import :: c_int, c_double, c_int64_t, c_funptr
         type(c_funptr), value :: res, jac, psol, rt
         integer(c_int), intent(in) :: neq, nrt
         real(c_double), intent(inout) :: t, y(*), yprime(*), rtol(*), atol(*), rwork(*), rpar(*)
         real(c_double), intent(in) :: tout
         integer(c_int), intent(inout) :: info(20), idid, iwork(*), ipar(*), jroot(*)
         integer(c_int64_t), intent(in) :: lrw, liw
      end subroutine

      ! ---- built-in device problem routines, used as addresses (c_funloc) ----
      function dkb_web_setpar(np, mx, my) bind(C, name="dkb_web_setpar") result(rc)
         import :: c_int
         integer(c_int), value :: np, mx, my
         integer(c_int) :: rc
      end function
      function dkb_web_cinit(c, cdot, pred_ic) bind(C, name="dkb_web_cinit") result(rc)
         import :: c_int, c_double
         real(c_double), intent(out) :: c(*), cdot(*)
         real(c_double), value :: pred_ic
         integer(c_int) :: rc
      end function
      subroutine dkb_web_res() bind(C, name="dkb_web_res")
      end subroutine
      subroutine dkb_web_jac() bind(C, name="dkb_web_jac")
      end subroutine
      subroutine dkb_web_psol() bind(C, name="dkb_web_psol")
      end subroutine
      subroutine dkb_web_rt() bind(C, name="dkb_web_rt")
      end subroutine

      ! ---- workspace of a host that keeps DASKR's own driver ----
      function dkb_alloc_workspace(neq, maxl, kmp, lenwp, leniwp) bind(C, name="dkb_alloc_workspace") result(rc)
         import :: c_int, c_int64_t
         integer(c_int64_t), value :: neq, lenwp, leniwp
         integer(c_int), value :: maxl, kmp
         integer(c_int) :: rc
      end function
      function dkb_workspace_get(which, ptr, ld) bind(C, name="dkb_workspace_get") result(rc)
         import :: c_int, c_int64_t, c_ptr
         integer(c_int), value :: which
         type(c_ptr), intent(out) :: ptr
         integer(c_int64_t), intent(out) :: ld
         integer(c_int) :: rc
      end function

      ! ---- the hot path: dslvk (src/daskr_new.f90:3036-3165) and ddwnrm (:828-868) on device arrays ----
      subroutine dslvk_dev(neq, y, t, ydot, savr, x, ewt, iwm, res, ires, psol, iersl, cj, epslin, sqrtn, rsqrtn, &
                           rhok, rpar, ipar) bind(C, name="dkb_dslvk")
         import :: c_int, c_double, c_ptr, c_funptr
         integer(c_int), intent(in) :: neq
         type(c_ptr), value :: y, ydot, savr, x, ewt
         real(c_double), intent(in) :: t, cj, epslin, sqrtn, rsqrtn
         integer(c_int), intent(inout) :: iwm(*), ires, iersl, ipar(*)
         type(c_funptr), value :: res, psol
         real(c_double), intent(inout) :: rhok, rpar(*)
      end subroutine
      function ddwnrm_dev(neq, v, rwt) bind(C, name="dkb_ddwnrm") result(r)
         import :: c_int, c_double, c_ptr
         integer(c_int), intent(in) :: neq
         type(c_ptr), value :: v, rwt
         real(c_double) :: r
      end function

      ! ---- the whole-vector statements of the host logic (csrc/vecapi.cu) ----
      function dkb_vcopy(n, x, y) bind(C, name="dkb_vcopy") result(rc)
         import :: c_int, c_int64_t, c_ptr
         integer(c_int64_t), value :: n
         type(c_ptr), value :: x, y
         integer(c_int) :: rc
      end function
      function dkb_vscale(n, a, x) bind(C, name="dkb_vscale") result(rc)
         import :: c_int, c_int64_t, c_double, c_ptr
         integer(c_int64_t), value :: n
         real(c_double), value :: a
         type(c_ptr), value :: x
         integer(c_int) :: rc
      end function
      function dkb_vscale_to(n, a, x, y) bind(C, name="dkb_vscale_to") result(rc)
         import :: c_int, c_int64_t, c_double, c_ptr
         integer(c_int64_t), value :: n
         real(c_double), value :: a
         type(c_ptr), value :: x, y
         integer(c_int) :: rc
      end function
      function dkb_vaxpy(n, a, x, y) bind(C, name="dkb_vaxpy") result(rc)
         import :: c_int, c_int64_t, c_double, c_ptr
         integer(c_int64_t), value :: n
         real(c_double), value :: a
         type(c_ptr), value :: x, y
         integer(c_int) :: rc
      end function
      function dkb_lincomb(n, a, x, b, y, z) bind(C, name="dkb_lincomb") result(rc)
         import :: c_int, c_int64_t, c_double, c_ptr
         integer(c_int64_t), value :: n
         real(c_double), value :: a, b
         type(c_ptr), value :: x, y, z
         integer(c_int) :: rc
      end function
      function dkb_ewt_set_invert(n, iwt, rtol, atol, y, wt, vt, id, ier) bind(C, name="dkb_ewt_set_invert") result(rc)
         import :: c_int, c_int64_t, c_double, c_ptr
         integer(c_int64_t), value :: n
         integer(c_int), value :: iwt
         real(c_double), intent(in) :: rtol(*), atol(*)      ! iwt = 1: pass device addresses through c_f_pointer-free wrappers
         type(c_ptr), value :: y, wt, vt, id
         integer(c_int), intent(out) :: ier
         integer(c_int) :: rc
      end function
      function dkb_mask_vt(n, id, wt, vt) bind(C, name="dkb_mask_vt") result(rc)
         import :: c_int, c_int64_t, c_ptr
         integer(c_int64_t), value :: n
         type(c_ptr), value :: id, wt, vt
         integer(c_int) :: rc
      end function
      function dkb_interp(n, t, tout, kold, phi, ldphi, psi, yout, ypout) bind(C, name="dkb_interp") result(rc)
         import :: c_int, c_int64_t, c_double, c_ptr
         integer(c_int64_t), value :: n, ldphi
         real(c_double), value :: t, tout
         integer(c_int), value :: kold
         type(c_ptr), value :: phi, yout, ypout
         real(c_double), intent(in) :: psi(*)
         integer(c_int) :: rc
      end function
      function dkb_predict(n, kp1, phi, ldphi, gama, y, yprime) bind(C, name="dkb_predict") result(rc)
         import :: c_int, c_int64_t, c_double, c_ptr
         integer(c_int64_t), value :: n, ldphi
         integer(c_int), value :: kp1
         type(c_ptr), value :: phi, y, yprime
         real(c_double), intent(in) :: gama(*)
         integer(c_int) :: rc
      end function
      function dkb_newton_update(n, delta, cj, y, e, yprime) bind(C, name="dkb_newton_update") result(rc)
         import :: c_int, c_int64_t, c_double, c_ptr
         integer(c_int64_t), value :: n
         type(c_ptr), value :: delta, y, e, yprime
         real(c_double), value :: cj
         integer(c_int) :: rc
      end function
      function dkb_phi_update(n, kp1, write_kp2, phi, ldphi, e) bind(C, name="dkb_phi_update") result(rc)
         import :: c_int, c_int64_t, c_ptr
         integer(c_int64_t), value :: n, ldphi
         integer(c_int), value :: kp1, write_kp2
         type(c_ptr), value :: phi, e
         integer(c_int) :: rc
      end function
      function dkb_error_norms(n, nn, e, rwt, phi_kp1, phi_k, out) bind(C, name="dkb_error_norms") result(rc)
         import :: c_int, c_int64_t, c_double, c_ptr
         integer(c_int64_t), value :: n
         integer(c_int), value :: nn
         type(c_ptr), value :: e, rwt, phi_kp1, phi_k
         real(c_double), intent(out) :: out(3)
         integer(c_int) :: rc
      end function

      ! ---- batched LINPACK (src/linpack.f90:325-520) on the reference's bd / ipbd layout ----
      function dkb_dgefa_batched(bd, ipbd, ns, nblk, info, info_first) bind(C, name="dkb_dgefa_batched") result(rc)
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: bd, ipbd, info
         integer(c_int), value :: ns
         integer(c_int64_t), value :: nblk
         integer(c_int64_t), intent(out) :: info_first
         integer(c_int) :: rc
      end function
      function dkb_dgesl_batched(bd, ipbd, ns, nblk, b) bind(C, name="dkb_dgesl_batched") result(rc)
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: bd, ipbd, b
         integer(c_int), value :: ns
         integer(c_int64_t), value :: nblk
         integer(c_int) :: rc
      end function

      ! ---- device memory for hosts without a CUDA Fortran binding ----
      function dkb_malloc(ptr, bytes) bind(C, name="dkb_malloc") result(rc)
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), intent(out) :: ptr
         integer(c_int64_t), value :: bytes
         integer(c_int) :: rc
      end function
      function dkb_free(ptr) bind(C, name="dkb_free") result(rc)
         import :: c_int, c_ptr
         type(c_ptr), value :: ptr
         integer(c_int) :: rc
      end function
      function dkb_memcpy_h2d(dst, src, bytes) bind(C, name="dkb_memcpy_h2d") result(rc)
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: dst, src
         integer(c_int64_t), value :: bytes
         integer(c_int) :: rc
      end function
      function dkb_memcpy_d2h(dst, src, bytes) bind(C, name="dkb_memcpy_d2h") result(rc)
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: dst, src
         integer(c_int64_t), value :: bytes
         integer(c_int) :: rc
      end function

      ! ---- checkpoint ----
      function dkb_state_bytes() bind(C, name="dkb_state_bytes") result(n)
         import :: c_int64_t
         integer(c_int64_t) :: n
      end function
      function dkb_state_export(buf, nbytes) bind(C, name="dkb_state_export") result(rc)
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: buf
         integer(c_int64_t), value :: nbytes
         integer(c_int) :: rc
      end function
      function dkb_state_import(buf, nbytes) bind(C, name="dkb_state_import") result(rc)
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: buf
         integer(c_int64_t), value :: nbytes
         integer(c_int) :: rc
      end function
      ! ---- library state, stream, errors ----
      function dkb_version() bind(C, name="dkb_version") result(v)
         import :: c_int
         integer(c_int) :: v
      end function
      !> text of the last error (a NUL-terminated C string: convert with c_f_pointer)
      function dkb_last_error() bind(C, name="dkb_last_error") result(msg)
         import :: c_ptr
         type(c_ptr) :: msg
      end function
      !> the cudaStream_t every launch of the library goes to (user callbacks enqueue there too)
      function dkb_stream() bind(C, name="dkb_stream") result(st)
         import :: c_ptr
         type(c_ptr) :: st
      end function
      function dkb_sync() bind(C, name="dkb_sync") result(rc)
         import :: c_int
         integer(c_int) :: rc
      end function
      !> DKB_OPT_ASYNC_OUT: block until the y / yprime of the last return have arrived in the caller's (pinned) arrays
      function dkb_output_wait() bind(C, name="dkb_output_wait") result(rc)
         import :: c_int
         integer(c_int) :: rc
      end function

      ! ---- food web on [0,1] x [0,ay]; heat problem (example/example_heat.f90) ----
      function dkb_web_setpar_ex(np, mx, my, ay) bind(C, name="dkb_web_setpar_ex") result(rc)
         import :: c_int, c_double
         integer(c_int), value :: np, mx, my
         real(c_double), value :: ay
         integer(c_int) :: rc
      end function
      function dkb_heat_setpar(m) bind(C, name="dkb_heat_setpar") result(rc)
         import :: c_int
         integer(c_int), value :: m
         integer(c_int) :: rc
      end function
      function dkb_heat_uinit(u, udot) bind(C, name="dkb_heat_uinit") result(rc)
         import :: c_int, c_double
         real(c_double), intent(out) :: u(*), udot(*)
         integer(c_int) :: rc
      end function
      !> res / jac / psol / rt of the heat example: pass c_funloc(dkb_heat_res) etc. to dkb_daskr (never called from Fortran)
      subroutine dkb_heat_res() bind(C, name="dkb_heat_res")
      end subroutine
      subroutine dkb_heat_jac() bind(C, name="dkb_heat_jac")
      end subroutine
      subroutine dkb_heat_psol() bind(C, name="dkb_heat_psol")
      end subroutine
      subroutine dkb_heat_rt() bind(C, name="dkb_heat_rt")
      end subroutine

      ! ---- ddot (extern/dblas.f:138) on device vectors ----
      function dkb_ddot(n, x, y, res) bind(C, name="dkb_ddot") result(rc)
         import :: c_int, c_int64_t, c_double, c_ptr
         integer(c_int64_t), value :: n
         type(c_ptr), value :: x, y
         real(c_double), intent(out) :: res
         integer(c_int) :: rc
      end function

      ! ---- the preconditioner's internal (tile-interleaved) block order and the reference's bd / ipbd ----
      function dkb_rbd_lenwp() bind(C, name="dkb_rbd_lenwp") result(n)
         import :: c_int64_t
         integer(c_int64_t) :: n
      end function
      function dkb_rbd_leniwp() bind(C, name="dkb_rbd_leniwp") result(n)
         import :: c_int64_t
         integer(c_int64_t) :: n
      end function
      function dkb_rbd_to_reference(rwp, iwp, bd, ipbd) bind(C, name="dkb_rbd_to_reference") result(rc)
         import :: c_int, c_ptr
         type(c_ptr), value :: rwp, iwp, bd, ipbd
         integer(c_int) :: rc
      end function
      function dkb_rbd_from_reference(bd, ipbd, rwp, iwp) bind(C, name="dkb_rbd_from_reference") result(rc)
         import :: c_int, c_ptr
         type(c_ptr), value :: bd, ipbd, rwp, iwp
         integer(c_int) :: rc
      end function

      ! ---- device memory helpers, measurement hooks ----
      function dkb_memcpy_d2d(dst, src, bytes) bind(C, name="dkb_memcpy_d2d") result(rc)
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: dst, src
         integer(c_int64_t), value :: bytes
         integer(c_int) :: rc
      end function
      function dkb_memset(dst, val, bytes) bind(C, name="dkb_memset") result(rc)
         import :: c_int, c_int64_t, c_ptr
         type(c_ptr), value :: dst
         integer(c_int), value :: val
         integer(c_int64_t), value :: bytes
         integer(c_int) :: rc
      end function
      function dkb_prof_enable(on) bind(C, name="dkb_prof_enable") result(rc)
         import :: c_int
         integer(c_int), value :: on
         integer(c_int) :: rc
      end function
      function dkb_prof_reset() bind(C, name="dkb_prof_reset") result(rc)
         import :: c_int
         integer(c_int) :: rc
      end function
      !> names: max_entries x 48 characters; ms, launches: max_entries each
      function dkb_prof_get(names, ms, launches, max_entries) bind(C, name="dkb_prof_get") result(n)
         import :: c_int, c_int64_t, c_double, c_char
         character(kind=c_char), intent(out) :: names(*)
         real(c_double), intent(out) :: ms(*)
         integer(c_int64_t), intent(out) :: launches(*)
         integer(c_int), value :: max_entries
         integer(c_int) :: n
      end function
      function dkb_launch_count() bind(C, name="dkb_launch_count") result(n)
         import :: c_int64_t
         integer(c_int64_t) :: n
      end function
      !> Krylov iterations since dkb_prof_reset by basis index ll = 1..n (dspigm's loop, src/daskr_new.f90:3304-3361)
      function dkb_krylov_stats(hist, n) bind(C, name="dkb_krylov_stats") result(rc)
         import :: c_int, c_int64_t
         integer(c_int64_t), intent(out) :: hist(*)
         integer(c_int), value :: n
         integer(c_int) :: rc
      end function
      !> dorth's conditional reorthogonalisation (src/daskr_new.f90:3575-3588): branch entered / corrections applied
      function dkb_reorth_stats(entered, corrections) bind(C, name="dkb_reorth_stats") result(rc)
         import :: c_int, c_int64_t
         integer(c_int64_t), intent(out) :: entered, corrections
         integer(c_int) :: rc
      end function
   end interface
end module daskr_b200
