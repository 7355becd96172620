!
! mod_snac_gpu_driver -- the glue between SNaC's driver (src/main.f90) and libsnacb200: what the six edits of
! integration/README.md call.  One MPI rank = one box = one GPU, exactly the reference's process model
! (src/initmpi.f90); the device is chosen by the rank's position on its node.
!
! UNVERIFIED SOURCE (no Fortran compiler in the build image of this repository).  The call sequence is the one
! snac_b200/sim.py + snac_b200/main.py execute through the same ABI, which is what the tests cover.
!
module mod_snac_gpu_driver
  use, intrinsic :: iso_c_binding
  use mpi
  use mod_types      , only: rp
  use mod_snac_gpu
  implicit none
  private
  public :: gpu_init, gpu_upload, gpu_download, gpu_step, gpu_bound_init, gpu_chkdt, gpu_chkdiv, gpu_finalize
  type(c_ptr), save :: h = c_null_ptr
  integer, save :: myrank = 0
  integer(c_int), allocatable, target, save :: owner(:)
contains
  subroutine check(ierr, what)
    integer(c_int), intent(in) :: ierr
    character(len=*), intent(in) :: what
    character(kind=c_char), pointer :: msg(:)
    type(c_ptr) :: cmsg
    integer :: n, ierr_mpi
    if(ierr == 0) return
    cmsg = snac_last_error(h)
    if(c_associated(cmsg)) then
      call c_f_pointer(cmsg, msg, [4096])
      n = 1
      do while(n < 4096 .and. msg(n) /= c_null_char)
        n = n + 1
      end do
      write(0,*) 'ERROR in '//what//': ', msg(1:n-1)
    end if
    call MPI_ABORT(MPI_COMM_WORLD, 1, ierr_mpi)   ! the reference's error convention (src/sanity.f90:191-198)
  end subroutine check
  !
  ! after initmpi / initgrid / bound_grid (src/main.f90:219-453): every rank describes EVERY box of the run to the
  ! library (the neighbour tables and the 1-D grids of the other ranks are gathered here), commits, and the
  ! CUDA-IPC blobs are exchanged so that halos and transposes go over NVLink.
  !
  subroutine gpu_init(myid, nrank, my_block, lo, hi, nb, is_bound, periods, fft_axis, impdiff, boussinesq, nscal, &
                      visc, bforce, gacc, beta, vol_all, alphai_s, ssource, is_sforced, scalf, &
                      hypre_tol, hypre_maxiter, hypre_solver_i, &
                      cbcvel, bcvel, cbcpre, bcpre, cbcscal, bcscal, dxc, dxf, dyc, dyf, dzc, dzf, dl, dlf)
    integer , intent(in) :: myid, nrank, my_block, lo(3), hi(3), nb(0:1,3), periods(3)
    logical , intent(in) :: is_bound(0:1,3)
    integer , intent(in) :: fft_axis, nscal, hypre_maxiter, hypre_solver_i
    logical , intent(in) :: impdiff, boussinesq
    real(rp), intent(in) :: visc, bforce(3), gacc(3), beta, vol_all, hypre_tol
    real(rp), intent(in) :: alphai_s(:), ssource(:), scalf(:)
    logical , intent(in) :: is_sforced(:)
    character(len=1), intent(in) :: cbcvel(0:1,3,3), cbcpre(0:1,3), cbcscal(:,:,:)   ! (0:1,3,nscal)
    real(rp), intent(in) :: bcvel(0:1,3,3), bcpre(0:1,3), bcscal(:,:,:)
    real(rp), intent(in), target :: dxc(lo(1)-1:), dxf(lo(1)-1:), dyc(lo(2)-1:), dyf(lo(2)-1:), dzc(lo(3)-1:), dzf(lo(3)-1:)
    real(rp), intent(in) :: dl(0:1,3), dlf(0:1,3)    ! d?c_g(lo_g-1), d?c_g(hi_g) ; d?f_g(lo_g), d?f_g(hi_g)
    type(snac_params) :: prm
    type(snac_box), allocatable :: boxes(:)
    type(snac_box) :: mine
    ! packed description of a box for MPI_ALLGATHER: integers and reals separately
    integer , allocatable :: ibuf(:,:)
    real(rp), allocatable :: rbuf(:,:)
    real(rp), allocatable, target :: grids(:,:,:)    ! (nmax+2, 6, nrank): dxc,dxf,dyc,dyf,dzc,dzf of every rank
    character(len=1), allocatable :: cbuf(:,:)
    integer :: ierr, r, d, ib, c, s, nmax, n(3), nloc, ndev, local_rank, comm_node
    integer(c_int) :: rc
    character(kind=c_char), allocatable, target :: blob(:), blobs(:)
    integer(c_size_t) :: nblob
    !
    myrank = myid
    call MPI_COMM_SPLIT_TYPE(MPI_COMM_WORLD, MPI_COMM_TYPE_SHARED, 0, MPI_INFO_NULL, comm_node, ierr)
    call MPI_COMM_RANK(comm_node, local_rank, ierr)
    allocate(owner(0:nrank-1))
    do r = 0, nrank-1
      owner(r) = r                                   ! one box per process
    end do
    prm%abi_version = SNAC_ABI_VERSION
    prm%device = local_rank
    prm%nrank = nrank; prm%nproc = nrank; prm%myproc = myid
    prm%rank_owner = c_loc(owner)
    prm%periods = periods
    prm%fft_axis = fft_axis
    prm%impdiff = merge(1, 0, impdiff); prm%boussinesq = merge(1, 0, boussinesq); prm%nscal = nscal
    prm%visc = visc; prm%bforce = bforce; prm%gacc = gacc; prm%beta = beta; prm%vol_all = vol_all
    prm%alphai = 0._rp; prm%ssource = 0._rp; prm%is_sforced = 0; prm%scalf = 0._rp
    do s = 1, nscal
      prm%alphai(s) = alphai_s(s); prm%ssource(s) = ssource(s)
      prm%is_sforced(s) = merge(1, 0, is_sforced(s)); prm%scalf(s) = scalf(s)
    end do
    prm%hypre_tol = hypre_tol; prm%hypre_maxiter = hypre_maxiter; prm%hypre_solver_i = hypre_solver_i
    prm%mg_npre = 0; prm%mg_npost = 0; prm%mg_nbottom = 0; prm%mg_generic = 0; prm%plane_solver = 0
    call check(snac_create(prm, h), 'snac_create')
    !
    ! gather the description of every box
    !
    n(:) = hi(:) - lo(:) + 1
    nloc = maxval(n) + 2
    call MPI_ALLREDUCE(nloc, nmax, 1, MPI_INTEGER, MPI_MAX, MPI_COMM_WORLD, ierr)
    allocate(ibuf(1+3+3+6+6, 0:nrank-1), rbuf(18+6+6+6+6*SNAC_MAX_SCAL, 0:nrank-1), cbuf(18+6+6*SNAC_MAX_SCAL, 0:nrank-1))
    allocate(grids(nmax, 6, 0:nrank-1))
    block
      integer  :: imine(19)
      real(rp) :: rmine(size(rbuf, 1)), gmine(nmax, 6)
      character(len=1) :: cmine(size(cbuf, 1))
      imine(1) = my_block - 1; imine(2:4) = lo; imine(5:7) = hi
      imine(8:13) = reshape(nb, [6]); imine(14:19) = merge(1, 0, reshape(is_bound, [6]))
      rmine = 0._rp; cmine = 'F'
      rmine(1:18) = reshape(bcvel, [18]); rmine(19:24) = reshape(bcpre, [6])
      rmine(25:30) = reshape(dl, [6]); rmine(31:36) = reshape(dlf, [6])
      cmine(1:18) = reshape(cbcvel, [18]); cmine(19:24) = reshape(cbcpre, [6])
      do s = 1, nscal
        rmine(36+6*(s-1)+1:36+6*s) = reshape(bcscal(:,:,s), [6])
        cmine(24+6*(s-1)+1:24+6*s) = reshape(cbcscal(:,:,s), [6])
      end do
      gmine = 1._rp
      gmine(1:n(1)+2, 1) = dxc(lo(1)-1:hi(1)+1); gmine(1:n(1)+2, 2) = dxf(lo(1)-1:hi(1)+1)
      gmine(1:n(2)+2, 3) = dyc(lo(2)-1:hi(2)+1); gmine(1:n(2)+2, 4) = dyf(lo(2)-1:hi(2)+1)
      gmine(1:n(3)+2, 5) = dzc(lo(3)-1:hi(3)+1); gmine(1:n(3)+2, 6) = dzf(lo(3)-1:hi(3)+1)
      call MPI_ALLGATHER(imine, size(imine), MPI_INTEGER, ibuf, size(imine), MPI_INTEGER, MPI_COMM_WORLD, ierr)
      call MPI_ALLGATHER(rmine, size(rmine), MPI_DOUBLE_PRECISION, rbuf, size(rmine), MPI_DOUBLE_PRECISION, MPI_COMM_WORLD, ierr)
      call MPI_ALLGATHER(cmine, size(cmine), MPI_CHARACTER, cbuf, size(cmine), MPI_CHARACTER, MPI_COMM_WORLD, ierr)
      call MPI_ALLGATHER(gmine, size(gmine), MPI_DOUBLE_PRECISION, grids, size(gmine), MPI_DOUBLE_PRECISION, MPI_COMM_WORLD, ierr)
    end block
    do r = 0, nrank-1
      mine%rank = r; mine%block = ibuf(1, r)
      mine%lo = ibuf(2:4, r); mine%hi = ibuf(5:7, r)
      mine%nb = reshape(ibuf(8:13, r), [2, 3])
      where(mine%nb == MPI_PROC_NULL) mine%nb = -1
      mine%is_bound = reshape(ibuf(14:19, r), [2, 3])
      mine%bcvel = reshape(rbuf(1:18, r), [2, 3, 3]); mine%bcpre = reshape(rbuf(19:24, r), [2, 3])
      mine%dl = reshape(rbuf(25:30, r), [2, 3]); mine%dlf = reshape(rbuf(31:36, r), [2, 3])
      mine%cbcvel = reshape(cbuf(1:18, r), [2, 3, 3]); mine%cbcpre = reshape(cbuf(19:24, r), [2, 3])
      mine%bcscal = 0._rp; mine%cbcscal = 'F'
      do s = 1, nscal
        mine%bcscal(:,:,s) = reshape(rbuf(36+6*(s-1)+1:36+6*s, r), [2, 3])
        mine%cbcscal(:,:,s) = reshape(cbuf(24+6*(s-1)+1:24+6*s, r), [2, 3])
      end do
      mine%dxc = c_loc(grids(1, 1, r)); mine%dxf = c_loc(grids(1, 2, r))
      mine%dyc = c_loc(grids(1, 3, r)); mine%dyf = c_loc(grids(1, 4, r))
      mine%dzc = c_loc(grids(1, 5, r)); mine%dzf = c_loc(grids(1, 6, r))
      call check(snac_add_box(h, mine), 'snac_add_box')
    end do
    call check(snac_commit(h), 'snac_commit')
    !
    ! peer memory: gather the fixed-size blobs in rank order
    !
    nblob = snac_comm_blob_size()
    allocate(blob(nblob), blobs(nblob*nrank))
    call check(snac_comm_export(h, c_loc(blob)), 'snac_comm_export')
    call MPI_ALLGATHER(blob, int(nblob), MPI_BYTE, blobs, int(nblob), MPI_BYTE, MPI_COMM_WORLD, ierr)
    call check(snac_comm_connect(h, c_loc(blobs), int(nrank, c_int)), 'snac_comm_connect')
    deallocate(blob, blobs, ibuf, rbuf, cbuf, grids)
  end subroutine gpu_init
  !
  ! after initflow / load_one (src/main.f90:527-556): fields to the device; arrays as allocated by the reference,
  ! u(lo(1)-1:hi(1)+1, lo(2)-1:hi(2)+1, lo(3)-1:hi(3)+1) (src/main.f90:226-231)
  !
  subroutine gpu_upload(u, v, w, p)
    real(rp), intent(in), contiguous :: u(:,:,:), v(:,:,:), w(:,:,:), p(:,:,:)
    call check(snac_upload(h, int(myrank, c_int), SNAC_FIELD_U, u), 'snac_upload(u)')
    call check(snac_upload(h, int(myrank, c_int), SNAC_FIELD_V, v), 'snac_upload(v)')
    call check(snac_upload(h, int(myrank, c_int), SNAC_FIELD_W, w), 'snac_upload(w)')
    call check(snac_upload(h, int(myrank, c_int), SNAC_FIELD_P, p), 'snac_upload(p)')
  end subroutine gpu_upload
  subroutine gpu_download(u, v, w, p)      ! before out1d/out2d/out3d and load_one('w') (src/main.f90:1073-1135)
    real(rp), intent(inout), contiguous :: u(:,:,:), v(:,:,:), w(:,:,:), p(:,:,:)
    call check(snac_download(h, int(myrank, c_int), SNAC_FIELD_U, u), 'snac_download(u)')
    call check(snac_download(h, int(myrank, c_int), SNAC_FIELD_V, v), 'snac_download(v)')
    call check(snac_download(h, int(myrank, c_int), SNAC_FIELD_W, w), 'snac_download(w)')
    call check(snac_download(h, int(myrank, c_int), SNAC_FIELD_P, p), 'snac_download(p)')
  end subroutine gpu_download
  subroutine gpu_bound_init()              ! bounduvw(.false.) + boundp(p) of src/main.f90:562-568
    call check(snac_bounduvw(h, 0_c_int), 'snac_bounduvw')
    call check(snac_boundp(h, SNAC_FIELD_P), 'snac_boundp')
  end subroutine gpu_bound_init
  subroutine gpu_step(dt)                  ! the whole `do irk=1,3 ... end do` (src/main.f90:906-1033)
    real(rp), intent(in) :: dt
    call check(snac_step(h, dt), 'snac_step')
  end subroutine gpu_step
  subroutine gpu_chkdt(dt_cfl)             ! src/main.f90:660,1051 (the result is global on every rank)
    real(rp), intent(out) :: dt_cfl
    call check(snac_chkdt(h, dt_cfl), 'snac_chkdt')
  end subroutine gpu_chkdt
  subroutine gpu_chkdiv(divtot, divmax)    ! src/main.f90:1061
    real(rp), intent(out) :: divtot, divmax
    call check(snac_chkdiv(h, divtot, divmax), 'snac_chkdiv')
  end subroutine gpu_chkdiv
  subroutine gpu_finalize()                ! in place of fftend / the HYPRE destroy calls (src/main.f90:1145-1174)
    integer(c_int) :: rc
    rc = snac_finalize(h)
    h = c_null_ptr
  end subroutine gpu_finalize
end module mod_snac_gpu_driver
