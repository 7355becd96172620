!
! mod_snac_gpu -- ISO_C_BINDING interface to libsnacb200 (include/snac_b200.h), written the way the reference
! binds its other native dependency (src/fftw.f90:9-28: one `bind(C)` interface block per C entry point).
!
! Drop this file into src/, add it to the Makefile's source list before main.f90, link with -lsnacb200 and
! apply integration/main_gpu.patch.  UNVERIFIED SOURCE: the build image of this repository has no Fortran
! compiler; the same ABI (struct layouts, argument order) is exercised by the ctypes binding snac_b200/lib.py
! and checked against the header by tests/test_abi.py.
!
! Conventions (see the header): boxes and blocks are 0-based in the ABI (rank = myid, block = my_block-1),
! lo/hi 1-based; the C arrays [idir][ib] / [ivel][idir][ib] have the memory order of the Fortran arrays
! (0:1,1:3) / (0:1,1:3,1:3), so cbcvel, bcvel, cbcpre, bcpre, nb, is_bound are passed as they are.
!
module mod_snac_gpu
  use, intrinsic :: iso_c_binding
  implicit none
  private
  public :: snac_params, snac_box
  public :: snac_create, snac_add_box, snac_commit, snac_comm_blob_size, snac_comm_export, snac_comm_connect, &
            snac_finalize, snac_last_error, snac_upload, snac_download, snac_rk_mom, snac_rk_scal, snac_bounduvw, &
            snac_boundp, snac_set_inflow, snac_inflow, snac_fillps, snac_updt_rhs, snac_fft, &
            snac_solve_n_helmholtz_2d, snac_solve_helmholtz, snac_correc, snac_updt_pressure, snac_chkdt, &
            snac_chkdiv, snac_out1d, snac_out2d, snac_step, snac_solver_stats, snac_solver_mean_iters, &
            snac_plane_solver, snac_sync, snac_download_async, snac_download_wait, snac_host_alloc, snac_host_free
  integer(c_int), parameter, public :: SNAC_ABI_VERSION = 2
  integer(c_int), parameter, public :: SNAC_MAX_SCAL = 8
  integer(c_int), parameter, public :: SNAC_FIELD_U = 0, SNAC_FIELD_V = 1, SNAC_FIELD_W = 2, SNAC_FIELD_P = 3, &
                                       SNAC_FIELD_PP = 4, SNAC_FIELD_RHS = 5, SNAC_FIELD_PO = 6, &
                                       SNAC_FIELD_DUDTRKO = 7, SNAC_FIELD_DVDTRKO = 8, SNAC_FIELD_DWDTRKO = 9, &
                                       SNAC_FIELD_RHS_HAT = 10, SNAC_FIELD_S0 = 16, SNAC_FIELD_DSDTRKO0 = 32
  !
  type, bind(C) :: snac_params            ! struct snac_params
    integer(c_int) :: abi_version, device, nrank, nproc, myproc
    type(c_ptr)    :: rank_owner          ! c_loc of an integer(c_int) array (0:nrank-1), or c_null_ptr
    integer(c_int) :: periods(3), fft_axis, impdiff, boussinesq, nscal
    real(c_double) :: visc, bforce(3), gacc(3), beta, vol_all
    real(c_double) :: alphai(SNAC_MAX_SCAL), ssource(SNAC_MAX_SCAL)
    integer(c_int) :: is_sforced(SNAC_MAX_SCAL)
    real(c_double) :: scalf(SNAC_MAX_SCAL), hypre_tol
    integer(c_int) :: hypre_maxiter, hypre_solver_i, mg_npre, mg_npost, mg_nbottom, mg_generic, plane_solver
  end type snac_params
  !
  type, bind(C) :: snac_box               ! struct snac_box
    integer(c_int) :: rank, block, lo(3), hi(3), nb(0:1,3), is_bound(0:1,3)
    character(kind=c_char) :: cbcvel(0:1,3,3)
    real(c_double) :: bcvel(0:1,3,3)
    character(kind=c_char) :: cbcpre(0:1,3)
    real(c_double) :: bcpre(0:1,3)
    character(kind=c_char) :: cbcscal(0:1,3,SNAC_MAX_SCAL)
    real(c_double) :: bcscal(0:1,3,SNAC_MAX_SCAL)
    type(c_ptr)    :: dxc, dxf, dyc, dyf, dzc, dzf   ! c_loc(dxc(lo(1)-1)) etc., AFTER bound_grid
    real(c_double) :: dl(0:1,3), dlf(0:1,3)
  end type snac_box
  !
  interface
    ! ---- life cycle ---------------------------------------------------------------------------
    integer(c_int) function snac_create(params, h) bind(C, name='snac_create')
      import :: c_int, c_ptr, snac_params
      type(snac_params), intent(in) :: params
      type(c_ptr), intent(out) :: h
    end function snac_create
    integer(c_int) function snac_add_box(h, box) bind(C, name='snac_add_box')
      import :: c_int, c_ptr, snac_box
      type(c_ptr), value :: h
      type(snac_box), intent(in) :: box
    end function snac_add_box
    integer(c_int) function snac_commit(h) bind(C, name='snac_commit')
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function snac_commit
    integer(c_size_t) function snac_comm_blob_size() bind(C, name='snac_comm_blob_size')
      import :: c_size_t
    end function snac_comm_blob_size
    integer(c_int) function snac_comm_export(h, blob) bind(C, name='snac_comm_export')
      import :: c_int, c_ptr
      type(c_ptr), value :: h, blob
    end function snac_comm_export
    integer(c_int) function snac_comm_connect(h, blobs, nblobs) bind(C, name='snac_comm_connect')
      import :: c_int, c_ptr
      type(c_ptr), value :: h, blobs
      integer(c_int), value :: nblobs
    end function snac_comm_connect
    integer(c_int) function snac_finalize(h) bind(C, name='snac_finalize')
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function snac_finalize
    type(c_ptr) function snac_last_error(h) bind(C, name='snac_last_error')   ! NUL-terminated C string
      import :: c_ptr
      type(c_ptr), value :: h
    end function snac_last_error
    ! ---- field transfer: host = u(lo(1)-1,lo(2)-1,lo(3)-1) of the reference's allocations (src/main.f90:226-231)
    integer(c_int) function snac_upload(h, rank, field, host) bind(C, name='snac_upload')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: rank, field
      real(c_double), intent(in) :: host(*)
    end function snac_upload
    integer(c_int) function snac_download(h, rank, field, host) bind(C, name='snac_download')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: rank, field
      real(c_double), intent(inout) :: host(*)
    end function snac_download
    ! outputs and checkpoints without stopping the time loop: `host` in pinned memory (snac_host_alloc + c_f_pointer)
    integer(c_int) function snac_download_async(h, rank, field, host, ticket) bind(C, name='snac_download_async')
      import :: c_int, c_ptr
      type(c_ptr), value :: h
      integer(c_int), value :: rank, field
      type(c_ptr), value :: host
      integer(c_int), intent(out) :: ticket
    end function snac_download_async
    integer(c_int) function snac_download_wait(h, ticket) bind(C, name='snac_download_wait')
      import :: c_int, c_ptr
      type(c_ptr), value :: h
      integer(c_int), value :: ticket
    end function snac_download_wait
    type(c_ptr) function snac_host_alloc(bytes) bind(C, name='snac_host_alloc')
      import :: c_ptr, c_size_t
      integer(c_size_t), value :: bytes
    end function snac_host_alloc
    subroutine snac_host_free(p) bind(C, name='snac_host_free')
      import :: c_ptr
      type(c_ptr), value :: p
    end subroutine snac_host_free
    ! ---- fine-grained entry points, 1:1 with the reference subroutines ---------------------------
    integer(c_int) function snac_rk_mom(h, irk, dt) bind(C, name='snac_rk_mom')                 ! src/rk.f90:7-128
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: irk
      real(c_double), value :: dt
    end function snac_rk_mom
    integer(c_int) function snac_rk_scal(h, irk, dt, iscal) bind(C, name='snac_rk_scal')        ! src/rk.f90:130-204
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: irk, iscal
      real(c_double), value :: dt
    end function snac_rk_scal
    integer(c_int) function snac_bounduvw(h, is_correc) bind(C, name='snac_bounduvw')           ! src/bound.f90:8-66
      import :: c_int, c_ptr
      type(c_ptr), value :: h
      integer(c_int), value :: is_correc
    end function snac_bounduvw
    integer(c_int) function snac_boundp(h, field) bind(C, name='snac_boundp')                   ! src/bound.f90:68-106
      import :: c_int, c_ptr
      type(c_ptr), value :: h
      integer(c_int), value :: field
    end function snac_boundp
    integer(c_int) function snac_set_inflow(h, rank, idir, ib, veln, velt1, velt2) bind(C, name='snac_set_inflow')
      import :: c_int, c_ptr, c_double                                                          ! src/main.f90:594-632
      type(c_ptr), value :: h
      integer(c_int), value :: rank, idir, ib      ! idir = 0..2, ib = 0 (lower) / 1 (upper)
      real(c_double), intent(in) :: veln(*), velt1(*), velt2(*)
    end function snac_set_inflow
    integer(c_int) function snac_inflow(h, is_correc) bind(C, name='snac_inflow')               ! src/bound.f90:438-477
      import :: c_int, c_ptr
      type(c_ptr), value :: h
      integer(c_int), value :: is_correc
    end function snac_inflow
    integer(c_int) function snac_fillps(h, dtrk) bind(C, name='snac_fillps')                    ! src/fillps.f90:7-54
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      real(c_double), value :: dtrk
    end function snac_fillps
    integer(c_int) function snac_updt_rhs(h) bind(C, name='snac_updt_rhs')                      ! src/bound.f90:543-585
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function snac_updt_rhs
    integer(c_int) function snac_fft(h, dir) bind(C, name='snac_fft')                           ! src/fft.f90:85-94
      import :: c_int, c_ptr
      type(c_ptr), value :: h
      integer(c_int), value :: dir                 ! +1 forward, -1 backward (times normfft, src/main.f90:1017-1020)
    end function snac_fft
    integer(c_int) function snac_solve_n_helmholtz_2d(h) bind(C, name='snac_solve_n_helmholtz_2d')   ! src/solver.f90:843-922
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function snac_solve_n_helmholtz_2d
    integer(c_int) function snac_solve_helmholtz(h) bind(C, name='snac_solve_helmholtz')       ! src/solver.f90:524-573
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function snac_solve_helmholtz
    integer(c_int) function snac_correc(h, dtrk) bind(C, name='snac_correc')                    ! src/correc.f90:7-48
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      real(c_double), value :: dtrk
    end function snac_correc
    integer(c_int) function snac_updt_pressure(h, alpha) bind(C, name='snac_updt_pressure')    ! src/updt_pressure.f90:7-47
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      real(c_double), value :: alpha
    end function snac_updt_pressure
    integer(c_int) function snac_chkdt(h, dt_cfl) bind(C, name='snac_chkdt')                    ! src/chkdt.f90:8-56
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      real(c_double), intent(out) :: dt_cfl
    end function snac_chkdt
    integer(c_int) function snac_chkdiv(h, divtot, divmax) bind(C, name='snac_chkdiv')         ! src/chkdiv.f90:8-44
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      real(c_double), intent(out) :: divtot, divmax
    end function snac_chkdiv
    integer(c_int) function snac_out1d(h, rank, field, idir, wa, wb, scale, p1d) bind(C, name='snac_out1d')   ! src/output.f90:34-133
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: rank, field, idir
      real(c_double), intent(in) :: wa(*), wb(*)
      real(c_double), value :: scale
      real(c_double), intent(out) :: p1d(*)
    end function snac_out1d
    integer(c_int) function snac_out2d(h, rank, field, idir, w, scale, p2d) bind(C, name='snac_out2d')       ! src/output.f90:135-225
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      integer(c_int), value :: rank, field, idir
      real(c_double), intent(in) :: w(*)
      real(c_double), value :: scale
      real(c_double), intent(out) :: p2d(*)
    end function snac_out2d
    ! ---- coarse-grained: the whole `do irk = 1,3` body (src/main.f90:906-1033) -------------------
    integer(c_int) function snac_step(h, dt) bind(C, name='snac_step')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      real(c_double), value :: dt
    end function snac_step
    ! ---- diagnostics -------------------------------------------------------------------------
    integer(c_int) function snac_solver_stats(h, iters, relres) bind(C, name='snac_solver_stats')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      integer(c_int), intent(out) :: iters
      real(c_double), intent(out) :: relres
    end function snac_solver_stats
    integer(c_int) function snac_solver_mean_iters(h, mean_iters) bind(C, name='snac_solver_mean_iters')
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: h
      real(c_double), intent(out) :: mean_iters
    end function snac_solver_mean_iters
    integer(c_int) function snac_plane_solver(h) bind(C, name='snac_plane_solver')
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function snac_plane_solver
    integer(c_int) function snac_sync(h) bind(C, name='snac_sync')
      import :: c_int, c_ptr
      type(c_ptr), value :: h
    end function snac_sync
  end interface
end module mod_snac_gpu
