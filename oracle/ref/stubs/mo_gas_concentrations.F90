! Stub of rte-rrtmgp's mo_gas_concentrations: ty_gas_concs with scalar / (nlay) / (ncol,nlay) concentrations and
! the broadcasting get_vmr the reference calls (src/gas_optics_ecckd.f90:343-351).  TEST INFRASTRUCTURE ONLY --
! written from the published interface (names lower-cased, profile broadcast over columns, size checks).
module mo_gas_concentrations
  use mo_rte_kind, only: wp
  implicit none
  private
  integer, parameter :: GAS_NOT_IN_LIST = -1

  type, private :: conc_field
    real(wp), dimension(:,:), allocatable :: conc    ! (1,1), (1,nlay) or (ncol,nlay)
  end type conc_field

  type, public :: ty_gas_concs
    character(len=32), dimension(:), allocatable :: gas_name
    type(conc_field), dimension(:), allocatable :: concs
    integer :: ncol = 0, nlay = 0
  contains
    procedure, public :: init
    procedure, private :: set_vmr_scalar
    procedure, private :: set_vmr_1d
    procedure, private :: set_vmr_2d
    generic, public :: set_vmr => set_vmr_scalar, set_vmr_1d, set_vmr_2d
    procedure, private :: get_vmr_1d
    procedure, private :: get_vmr_2d
    generic, public :: get_vmr => get_vmr_1d, get_vmr_2d
    procedure, public :: get_num_gases
    procedure, public :: get_gas_names
  end type ty_gas_concs

contains

  pure function lower_case(s) result(t)
    character(len=*), intent(in) :: s
    character(len=len(s)) :: t
    integer :: i, c
    t = s
    do i = 1, len(s)
      c = iachar(s(i:i))
      if (c >= iachar('A') .and. c <= iachar('Z')) t(i:i) = achar(c + 32)
    end do
  end function lower_case

  function init(this, gas_names) result(error_msg)
    class(ty_gas_concs), intent(inout) :: this
    character(len=*), dimension(:), intent(in) :: gas_names
    character(len=128) :: error_msg
    integer :: i
    error_msg = ""
    if (allocated(this%gas_name)) deallocate(this%gas_name)
    if (allocated(this%concs)) deallocate(this%concs)
    allocate(this%gas_name(size(gas_names)), this%concs(size(gas_names)))
    do i = 1, size(gas_names)
      this%gas_name(i) = lower_case(trim(gas_names(i)))
    end do
    this%ncol = 0
    this%nlay = 0
  end function init

  pure function find_gas(this, gas) result(idx)
    class(ty_gas_concs), intent(in) :: this
    character(len=*), intent(in) :: gas
    integer :: idx, i
    idx = GAS_NOT_IN_LIST
    if (.not. allocated(this%gas_name)) return
    do i = 1, size(this%gas_name)
      if (trim(this%gas_name(i)) == trim(lower_case(gas))) then
        idx = i
        return
      end if
    end do
  end function find_gas

  function set_vmr_scalar(this, gas, w) result(error_msg)
    class(ty_gas_concs), intent(inout) :: this
    character(len=*), intent(in) :: gas
    real(wp), intent(in) :: w
    character(len=128) :: error_msg
    integer :: idx
    error_msg = ""
    idx = find_gas(this, gas)
    if (idx == GAS_NOT_IN_LIST) then
      error_msg = "ty_gas_concs%set_vmr(): trying to set " // trim(gas) // " but name not provided at initialization"
      return
    end if
    if (allocated(this%concs(idx)%conc)) deallocate(this%concs(idx)%conc)
    allocate(this%concs(idx)%conc(1, 1))
    this%concs(idx)%conc(1, 1) = w
  end function set_vmr_scalar

  function set_vmr_1d(this, gas, w) result(error_msg)
    class(ty_gas_concs), intent(inout) :: this
    character(len=*), intent(in) :: gas
    real(wp), dimension(:), intent(in) :: w
    character(len=128) :: error_msg
    integer :: idx
    error_msg = ""
    idx = find_gas(this, gas)
    if (idx == GAS_NOT_IN_LIST) then
      error_msg = "ty_gas_concs%set_vmr(): trying to set " // trim(gas) // " but name not provided at initialization"
      return
    end if
    if (this%nlay > 0 .and. size(w) /= this%nlay) then
      error_msg = "ty_gas_concs%set_vmr: different dimension (nlay)"
      return
    end if
    this%nlay = size(w)
    if (allocated(this%concs(idx)%conc)) deallocate(this%concs(idx)%conc)
    allocate(this%concs(idx)%conc(1, size(w)))
    this%concs(idx)%conc(1, :) = w(:)
  end function set_vmr_1d

  function set_vmr_2d(this, gas, w) result(error_msg)
    class(ty_gas_concs), intent(inout) :: this
    character(len=*), intent(in) :: gas
    real(wp), dimension(:,:), intent(in) :: w
    character(len=128) :: error_msg
    integer :: idx
    error_msg = ""
    idx = find_gas(this, gas)
    if (idx == GAS_NOT_IN_LIST) then
      error_msg = "ty_gas_concs%set_vmr(): trying to set " // trim(gas) // " but name not provided at initialization"
      return
    end if
    if (this%ncol > 0 .and. size(w, 1) /= this%ncol) then
      error_msg = "ty_gas_concs%set_vmr: different dimension (ncol)"
      return
    end if
    if (this%nlay > 0 .and. size(w, 2) /= this%nlay) then
      error_msg = "ty_gas_concs%set_vmr: different dimension (nlay)"
      return
    end if
    this%ncol = size(w, 1)
    this%nlay = size(w, 2)
    if (allocated(this%concs(idx)%conc)) deallocate(this%concs(idx)%conc)
    allocate(this%concs(idx)%conc(size(w, 1), size(w, 2)))
    this%concs(idx)%conc(:,:) = w(:,:)
  end function set_vmr_2d

  function get_vmr_1d(this, gas, array) result(error_msg)
    class(ty_gas_concs), intent(in) :: this
    character(len=*), intent(in) :: gas
    real(wp), dimension(:), intent(out) :: array
    character(len=128) :: error_msg
    integer :: idx
    error_msg = ""
    idx = find_gas(this, gas)
    if (idx == GAS_NOT_IN_LIST) then
      error_msg = "ty_gas_concs%get_vmr; gas " // trim(gas) // " not found"
      array(:) = 0._wp
      return
    end if
    if (.not. allocated(this%concs(idx)%conc)) then
      error_msg = "ty_gas_concs%get_vmr; gas " // trim(gas) // " not found"
      array(:) = 0._wp
      return
    end if
    if (size(this%concs(idx)%conc, 1) > 1) then
      error_msg = "ty_gas_concs%get_vmr; gas " // trim(gas) // " requested as 1D but stored as 2D"
    else if (size(this%concs(idx)%conc, 2) > 1 .and. size(this%concs(idx)%conc, 2) /= size(array)) then
      error_msg = "ty_gas_concs%get_vmr; gas " // trim(gas) // " array is wrong size (nlay)"
    end if
    if (error_msg /= "") return
    if (size(this%concs(idx)%conc, 2) > 1) then
      array(:) = this%concs(idx)%conc(1, :)
    else
      array(:) = this%concs(idx)%conc(1, 1)
    end if
  end function get_vmr_1d

  function get_vmr_2d(this, gas, array) result(error_msg)
    class(ty_gas_concs), intent(in) :: this
    character(len=*), intent(in) :: gas
    real(wp), dimension(:,:), intent(out) :: array
    character(len=128) :: error_msg
    integer :: idx, icol, ilay
    error_msg = ""
    idx = find_gas(this, gas)
    if (idx == GAS_NOT_IN_LIST) then
      error_msg = "ty_gas_concs%get_vmr; gas " // trim(gas) // " not found"
      array(:,:) = 0._wp
      return
    end if
    if (.not. allocated(this%concs(idx)%conc)) then
      error_msg = "ty_gas_concs%get_vmr; gas " // trim(gas) // " not found"
      array(:,:) = 0._wp
      return
    end if
    if (size(this%concs(idx)%conc, 1) > 1 .and. size(this%concs(idx)%conc, 1) /= size(array, 1)) then
      error_msg = "ty_gas_concs%get_vmr; gas " // trim(gas) // " array is wrong size (ncol)"
    end if
    if (size(this%concs(idx)%conc, 2) > 1 .and. size(this%concs(idx)%conc, 2) /= size(array, 2)) then
      error_msg = "ty_gas_concs%get_vmr; gas " // trim(gas) // " array is wrong size (nlay)"
    end if
    if (error_msg /= "") return
    if (size(this%concs(idx)%conc, 1) > 1) then            ! (ncol, nlay) as is
      array(:,:) = this%concs(idx)%conc(:,:)
    else if (size(this%concs(idx)%conc, 2) > 1) then        ! (nlay) profile on every column
      do ilay = 1, size(array, 2)
        do icol = 1, size(array, 1)
          array(icol, ilay) = this%concs(idx)%conc(1, ilay)
        end do
      end do
    else                                                     ! scalar everywhere
      array(:,:) = this%concs(idx)%conc(1, 1)
    end if
  end function get_vmr_2d

  pure function get_num_gases(this)
    class(ty_gas_concs), intent(in) :: this
    integer :: get_num_gases
    get_num_gases = 0
    if (allocated(this%gas_name)) get_num_gases = size(this%gas_name)
  end function get_num_gases

  pure function get_gas_names(this)
    class(ty_gas_concs), intent(in) :: this
    character(len=32), dimension(this%get_num_gases()) :: get_gas_names
    get_gas_names(:) = this%gas_name(:)
  end function get_gas_names

end module mo_gas_concentrations
