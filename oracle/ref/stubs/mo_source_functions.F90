! Stub of rte-rrtmgp's mo_source_functions (<= v1.6: lev_source_inc / lev_source_dec, which the reference
! writes at src/gas_optics_ecckd.f90:419-424).  TEST INFRASTRUCTURE ONLY.
module mo_source_functions
  use mo_rte_kind, only: wp
  use mo_optical_props, only: ty_optical_props
  implicit none
  private

  type, extends(ty_optical_props), public :: ty_source_func_lw
    real(wp), allocatable, dimension(:,:,:) :: lay_source, lev_source_inc, lev_source_dec  ! (ncol, nlay, ngpt)
    real(wp), allocatable, dimension(:,:) :: sfc_source, sfc_source_Jac                    ! (ncol, ngpt)
  contains
    procedure, public :: alloc => alloc_lw
  end type ty_source_func_lw

contains

  function alloc_lw(this, ncol, nlay, ngpt) result(err_message)
    class(ty_source_func_lw), intent(inout) :: this
    integer, intent(in) :: ncol, nlay, ngpt
    character(len=128) :: err_message
    err_message = ""
    if (allocated(this%lay_source)) deallocate(this%lay_source)
    if (allocated(this%lev_source_inc)) deallocate(this%lev_source_inc)
    if (allocated(this%lev_source_dec)) deallocate(this%lev_source_dec)
    if (allocated(this%sfc_source)) deallocate(this%sfc_source)
    if (allocated(this%sfc_source_Jac)) deallocate(this%sfc_source_Jac)
    allocate(this%lay_source(ncol, nlay, ngpt), this%lev_source_inc(ncol, nlay, ngpt), this%lev_source_dec(ncol, nlay, ngpt))
    allocate(this%sfc_source(ncol, ngpt), this%sfc_source_Jac(ncol, ngpt))
    this%lay_source(:,:,:) = 0._wp
    this%lev_source_inc(:,:,:) = 0._wp
    this%lev_source_dec(:,:,:) = 0._wp
    this%sfc_source(:,:) = 0._wp
    this%sfc_source_Jac(:,:) = 0._wp
  end function alloc_lw

end module mo_source_functions
