! Stub of rte-rrtmgp's mo_optical_props: just the parts the ecckd gas optics touches
! (src/gas_optics_ecckd.f90:5, :344-347, :370, :456-464; example/rfmip-rad-irf/mo_load_coefficients.F90:74).
! TEST INFRASTRUCTURE ONLY -- written from the published rte-rrtmgp interface, not copied from it.
module mo_optical_props
  use mo_rte_kind, only: wp
  implicit none
  private

  !> spectral discretisation: bands and the g-points they hold
  type, public :: ty_optical_props
    integer, dimension(:,:), allocatable :: band2gpt       ! (2, nband)
    real(wp), dimension(:,:), allocatable :: band_lims_wvn ! (2, nband)
    character(len=32) :: name = ""
  contains
    procedure, public :: init => init_base
    procedure, public :: get_nband
    procedure, public :: get_ngpt
    procedure, public :: finalize => finalize_base
  end type ty_optical_props

  !> optical properties as arrays (ncol, nlay, ngpt)
  type, extends(ty_optical_props), abstract, public :: ty_optical_props_arry
    real(wp), dimension(:,:,:), allocatable :: tau
  end type ty_optical_props_arry

  type, extends(ty_optical_props_arry), public :: ty_optical_props_1scl
  contains
    procedure, public :: alloc_1scl
  end type ty_optical_props_1scl

  type, extends(ty_optical_props_arry), public :: ty_optical_props_2str
    real(wp), dimension(:,:,:), allocatable :: ssa, g
  contains
    procedure, public :: alloc_2str
  end type ty_optical_props_2str

contains

  function init_base(this, band_lims_wvn, band_lims_gpt, name) result(err_message)
    class(ty_optical_props), intent(inout) :: this
    real(wp), dimension(:,:), intent(in) :: band_lims_wvn
    integer, dimension(:,:), intent(in), optional :: band_lims_gpt
    character(len=*), intent(in), optional :: name
    character(len=128) :: err_message
    integer :: ib
    err_message = ""
    if (size(band_lims_wvn, 1) /= 2) then
      err_message = "optical_props%init(): band_lims_wvn 1st dim should be 2"
      return
    end if
    if (allocated(this%band_lims_wvn)) deallocate(this%band_lims_wvn)
    if (allocated(this%band2gpt)) deallocate(this%band2gpt)
    allocate(this%band_lims_wvn(2, size(band_lims_wvn, 2)), this%band2gpt(2, size(band_lims_wvn, 2)))
    this%band_lims_wvn(:,:) = band_lims_wvn(:,:)
    if (present(band_lims_gpt)) then
      this%band2gpt(:,:) = band_lims_gpt(:,:)
    else
      do ib = 1, size(band_lims_wvn, 2)
        this%band2gpt(:,ib) = ib
      end do
    end if
    if (present(name)) this%name = trim(name)
  end function init_base

  subroutine finalize_base(this)
    class(ty_optical_props), intent(inout) :: this
    if (allocated(this%band2gpt)) deallocate(this%band2gpt)
    if (allocated(this%band_lims_wvn)) deallocate(this%band_lims_wvn)
    this%name = ""
  end subroutine finalize_base

  pure function get_nband(this)
    class(ty_optical_props), intent(in) :: this
    integer :: get_nband
    get_nband = 0
    if (allocated(this%band2gpt)) get_nband = size(this%band2gpt, 2)
  end function get_nband

  pure function get_ngpt(this)
    class(ty_optical_props), intent(in) :: this
    integer :: get_ngpt
    get_ngpt = 0
    if (allocated(this%band2gpt)) get_ngpt = maxval(this%band2gpt)
  end function get_ngpt

  function alloc_1scl(this, ncol, nlay, ngpt) result(err_message)
    class(ty_optical_props_1scl), intent(inout) :: this
    integer, intent(in) :: ncol, nlay, ngpt
    character(len=128) :: err_message
    err_message = ""
    if (allocated(this%tau)) deallocate(this%tau)
    allocate(this%tau(ncol, nlay, ngpt))
    this%tau(:,:,:) = 0._wp
  end function alloc_1scl

  function alloc_2str(this, ncol, nlay, ngpt) result(err_message)
    class(ty_optical_props_2str), intent(inout) :: this
    integer, intent(in) :: ncol, nlay, ngpt
    character(len=128) :: err_message
    err_message = ""
    if (allocated(this%tau)) deallocate(this%tau)
    if (allocated(this%ssa)) deallocate(this%ssa)
    if (allocated(this%g)) deallocate(this%g)
    allocate(this%tau(ncol, nlay, ngpt), this%ssa(ncol, nlay, ngpt), this%g(ncol, nlay, ngpt))
    this%tau(:,:,:) = 0._wp
    this%ssa(:,:,:) = 0._wp
    this%g(:,:,:) = 0._wp
  end function alloc_2str

end module mo_optical_props
