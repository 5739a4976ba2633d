!> @file gas_optics_ecckd_b200_shim.f90
!!
!! Drop-in replacement of /root/reference/src/gas_optics_ecckd.f90 for the gas-optics hot
!! path: the module name, the public type ty_gas_optics_ecckd with its public components, the
!! AbsorptionTable helper type, the concentration-dependence codes and every type-bound
!! procedure keep the reference's names, argument lists and error strings, so the reference's
!! own loader (example/rfmip-rad-irf/mo_load_coefficients.F90) and drivers compile against it
!! unchanged.  What changes is the body of gas_optics_int / gas_optics_ext: instead of the
!! Fortran loop nests (:64-241, :245-289, :293-319 of the reference) they hand the caller's
!! arrays to the C ABI of libecckd_b200.so (include/ecckd_b200.h), which runs sm_100a CUDA
!! kernels.  There is no CPU fallback: if the library or a device is missing the error string
!! says so and the driver's stop_on_err aborts.
!!
!! The device-resident tables are packed once, on the first gas_optics call (or explicitly
!! through finalize_load), from the components the loader filled.
!!
!! NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Fortran compiler and no
!! rte-rrtmgp.  INTEGRATION.md gives the build line; tests/ exercise the very same C entry
!! points through ctypes with the same argument meaning.
module gas_optics_ecckd
use, intrinsic :: iso_c_binding
use, intrinsic :: iso_fortran_env, only: error_unit
use mo_gas_concentrations, only: ty_gas_concs
use mo_gas_optics, only: ty_gas_optics
use mo_optical_props, only: ty_optical_props_arry, ty_optical_props_2str
use mo_rte_kind, only: wp
use mo_source_functions, only: ty_source_func_lw
implicit none
private


!> @brief Helper class used to hold the table of each gas (reference :13-19, unchanged).
type, public :: AbsorptionTable
  real(kind=wp), dimension(:,:,:,:), allocatable :: coefficient
  logical :: composite_only
  integer :: concentration_dependence_code
  real(kind=wp), dimension(:), allocatable :: mole_fraction
  real(kind=wp) :: reference_mole_fraction
end type AbsorptionTable


!> @brief ecckd gas optics (reference :23-48) plus the opaque device handle.
type, extends(ty_gas_optics), public :: ty_gas_optics_ecckd
  type(AbsorptionTable), dimension(16) :: absorption
  character(len=32), dimension(16) :: gas
  real(kind=wp), dimension(:,:), allocatable :: gpoint_fraction
  real(kind=wp), dimension(:), allocatable :: log_pressure
  integer :: num_composite_gases
  integer :: num_gases
  real(kind=wp), dimension(:,:), allocatable :: planck_function
  real(kind=wp), dimension(:), allocatable :: rayleigh_molar_scattering_coeff
  logical :: shortwave
  real(kind=wp), dimension(:), allocatable :: solar_irradiance
  real(kind=wp), dimension(:,:), allocatable :: temperature
  real(kind=wp), dimension(:), allocatable :: temperature_planck
  real(kind=wp) :: total_solar_irradiance
  type(c_ptr) :: device_model = c_null_ptr   !< ecckd_model*, created lazily (tables packed once)
  contains
  procedure, public :: source_is_internal
  procedure, public :: source_is_external
  procedure, public :: get_ngas
  procedure, public :: get_gases
  procedure, public :: get_press_min
  procedure, public :: get_press_max
  procedure, public :: get_temp_min
  procedure, public :: get_temp_max
  procedure, public :: gas_optics_int
  procedure, public :: gas_optics_ext
  procedure, public :: finalize_load
  procedure, public :: release_device
end type ty_gas_optics_ecckd


integer, parameter, public :: none_ = 0
integer, parameter, public :: linear = 1
integer, parameter, public :: look_up_table = 2
integer, parameter, public :: relative_linear = 3


! ---- mirrors of the C structs in include/ecckd_b200.h ------------------------------------
type, bind(c) :: ecckd_gas_table
  character(kind=c_char) :: name(32)
  integer(c_int32_t) :: composite_only
  integer(c_int32_t) :: concentration_dependence_code
  integer(c_int32_t) :: n_mole_fraction
  integer(c_int32_t) :: reserved
  real(c_double) :: reference_mole_fraction
  type(c_ptr) :: mole_fraction
  type(c_ptr) :: coefficient
end type ecckd_gas_table

type, bind(c) :: ecckd_host_tables
  integer(c_int32_t) :: shortwave, ngpt, npress, ntemp, nplanck, num_gases
  type(c_ptr) :: log_pressure, temperature, temperature_planck, planck_function
  type(c_ptr) :: rayleigh_molar_scattering_coeff, solar_irradiance
  type(c_ptr) :: gas
  integer(c_int32_t) :: nband, reserved
  type(c_ptr) :: band_lims_wvn, band2gpt
end type ecckd_host_tables

type, bind(c) :: ecckd_vmr_view
  type(c_ptr) :: ptr
  integer(c_int64_t) :: col_stride, lay_stride
end type ecckd_vmr_view


interface
  integer(c_int) function ecckd_model_create(tables, ndev, devices, model, err) bind(c, name="ecckd_model_create")
    import :: c_int, c_ptr, c_char, ecckd_host_tables
    type(ecckd_host_tables), intent(in) :: tables
    integer(c_int), value :: ndev
    type(c_ptr), value :: devices
    type(c_ptr), intent(out) :: model
    character(kind=c_char) :: err(128)
  end function
  subroutine ecckd_model_destroy(model) bind(c, name="ecckd_model_destroy")
    import :: c_ptr
    type(c_ptr), value :: model
  end subroutine
  integer(c_int) function ecckd_gas_optics_lw_host_f64(model, ncol, nlay, plev, tlay, tlev, tsfc, ngas, names, vmr, &
      tau, lay_source, lev_source_inc, lev_source_dec, sfc_source, block_cols, err) bind(c, name="ecckd_gas_optics_lw_host_f64")
    import :: c_int, c_ptr, c_char
    type(c_ptr), value :: model, plev, tlay, tlev, tsfc, names, vmr, tau, lay_source, lev_source_inc, lev_source_dec, sfc_source
    integer(c_int), value :: ncol, nlay, ngas, block_cols
    character(kind=c_char) :: err(128)
  end function
  integer(c_int) function ecckd_gas_optics_sw_host_f64(model, ncol, nlay, plev, tlay, ngas, names, vmr, &
      tau, ssa, g, toa_src, block_cols, err) bind(c, name="ecckd_gas_optics_sw_host_f64")
    import :: c_int, c_ptr, c_char
    type(c_ptr), value :: model, plev, tlay, names, vmr, tau, ssa, g, toa_src
    integer(c_int), value :: ncol, nlay, ngas, block_cols
    character(kind=c_char) :: err(128)
  end function
end interface


contains


!> @brief Pack the tables the loader filled and create the device model (idempotent).
function finalize_load(this) result(error_msg)
  class(ty_gas_optics_ecckd), intent(inout), target :: this
  character(len=128) :: error_msg

  type(ecckd_host_tables) :: t
  type(ecckd_gas_table), dimension(16), target :: g
  character(kind=c_char), dimension(128) :: err
  integer :: i, k, rc

  error_msg = ""
  if (c_associated(this%device_model)) return
  t%shortwave = merge(1, 0, this%shortwave)
  t%ngpt = size(this%gpoint_fraction, 2)
  t%npress = size(this%log_pressure)
  t%ntemp = size(this%temperature, 2)
  t%nplanck = 0
  t%num_gases = this%num_gases
  t%log_pressure = c_loc(this%log_pressure)
  t%temperature = c_loc(this%temperature)
  t%temperature_planck = c_null_ptr; t%planck_function = c_null_ptr
  t%rayleigh_molar_scattering_coeff = c_null_ptr; t%solar_irradiance = c_null_ptr
  if (allocated(this%temperature_planck)) then
    t%nplanck = size(this%temperature_planck)
    t%temperature_planck = c_loc(this%temperature_planck)
    t%planck_function = c_loc(this%planck_function)
  endif
  if (allocated(this%solar_irradiance)) then
    t%solar_irradiance = c_loc(this%solar_irradiance)
    t%rayleigh_molar_scattering_coeff = c_loc(this%rayleigh_molar_scattering_coeff)
  endif
  do i = 1, this%num_gases
    do k = 1, 32
      g(i)%name(k) = this%gas(i)(k:k)
    enddo
    g(i)%composite_only = merge(1, 0, this%absorption(i)%composite_only)
    g(i)%concentration_dependence_code = this%absorption(i)%concentration_dependence_code
    g(i)%n_mole_fraction = 0
    g(i)%mole_fraction = c_null_ptr
    if (allocated(this%absorption(i)%mole_fraction)) then
      g(i)%n_mole_fraction = size(this%absorption(i)%mole_fraction)
      g(i)%mole_fraction = c_loc(this%absorption(i)%mole_fraction)
    endif
    g(i)%reserved = 0
    g(i)%reference_mole_fraction = this%absorption(i)%reference_mole_fraction
    g(i)%coefficient = c_loc(this%absorption(i)%coefficient)
  enddo
  t%gas = c_loc(g)
  t%nband = 0; t%reserved = 0
  t%band_lims_wvn = c_null_ptr; t%band2gpt = c_null_ptr      ! kept by the inherited ty_optical_props%init
  rc = ecckd_model_create(t, 1, c_null_ptr, this%device_model, err)
  if (rc .ne. 0) call c_to_f_string(err, error_msg)
end function finalize_load


!> @brief Free the device model.
subroutine release_device(this)
  class(ty_gas_optics_ecckd), intent(inout) :: this
  if (c_associated(this%device_model)) call ecckd_model_destroy(this%device_model)
  this%device_model = c_null_ptr
end subroutine release_device


!> @brief Gather the caller's ty_gas_concs into names[] + ecckd_vmr_view[] (reference :340-357).
!! get_vmr materialises (ncol,nlay) exactly as the reference does (:343,351); a field that is
!! constant is handed over as a scalar view so that the library can pre-mix it (its fast path).
function marshal_gases(gas_desc, ncol, nlay, names, name_ptrs, views, vmr, ngas) result(error_msg)
  type(ty_gas_concs), intent(in) :: gas_desc
  integer, intent(in) :: ncol, nlay
  character(kind=c_char, len=33), dimension(:), allocatable, target, intent(out) :: names
  type(c_ptr), dimension(:), allocatable, target, intent(out) :: name_ptrs
  type(ecckd_vmr_view), dimension(:), allocatable, target, intent(out) :: views
  real(kind=wp), dimension(:,:,:), allocatable, target, intent(out) :: vmr
  integer, intent(out) :: ngas
  character(len=128) :: error_msg

  character(len=32), dimension(:), allocatable :: gas_names
  integer :: j

  error_msg = ""
  ngas = gas_desc%get_num_gases()
  allocate(gas_names(ngas), names(ngas), name_ptrs(ngas), views(ngas), vmr(ncol, nlay, ngas))
  gas_names = gas_desc%get_gas_names()
  do j = 1, ngas
    names(j) = trim(gas_names(j))//c_null_char
    name_ptrs(j) = c_loc(names(j))
    error_msg = gas_desc%get_vmr(gas_names(j), vmr(:,:,j))
    if (trim(error_msg) .ne. "") return
    views(j)%ptr = c_loc(vmr(1,1,j))
    if (minval(vmr(:,:,j)) .eq. maxval(vmr(:,:,j))) then
      views(j)%col_stride = 0; views(j)%lay_stride = 0
    else
      views(j)%col_stride = 1; views(j)%lay_stride = ncol
    endif
  enddo
end function marshal_gases


!> @brief Compute gas optical depth and Planck source functions (reference :381-426).
function gas_optics_int(this, play, plev, tlay, tsfc, gas_desc, &
                        optical_props, sources, col_dry, tlev) &
  result(error_msg)

  class(ty_gas_optics_ecckd), intent(in) :: this
  real(kind=wp), dimension(:,:), intent(in) :: play
  real(kind=wp), dimension(:,:), intent(in) :: plev
  real(kind=wp), dimension(:,:), intent(in) :: tlay
  real(kind=wp), dimension(:), intent(in) :: tsfc
  type(ty_gas_concs), intent(in) :: gas_desc
  class(ty_optical_props_arry), intent(inout) :: optical_props
  class(ty_source_func_lw), intent(inout) :: sources
  character(len=128) :: error_msg
  real(kind=wp), dimension(:,:), intent(in), target, optional :: col_dry
  real(kind=wp), dimension(:,:), intent(in), target, optional :: tlev

  character(kind=c_char, len=33), dimension(:), allocatable, target :: names
  type(c_ptr), dimension(:), allocatable, target :: name_ptrs
  type(ecckd_vmr_view), dimension(:), allocatable, target :: views
  real(kind=wp), dimension(:,:,:), allocatable, target :: vmr
  real(kind=wp), dimension(:,:), allocatable, target :: plev_c, tlay_c, tlev_c
  real(kind=wp), dimension(:), allocatable, target :: tsfc_c
  character(kind=c_char), dimension(128) :: err
  type(c_ptr) :: tlev_ptr
  integer :: ncol, nlay, ngas, rc

  if (.not. c_associated(this%device_model)) then
    error_msg = "ecckd: call finalize_load() after load_and_init (device tables not packed)"
    return
  endif
  ncol = size(tlay, 1); nlay = size(tlay, 2)
  error_msg = marshal_gases(gas_desc, ncol, nlay, names, name_ptrs, views, vmr, ngas)
  if (len_trim(error_msg) .gt. 0) return
  plev_c = plev; tlay_c = tlay; tsfc_c = tsfc            ! contiguous copies (assumed-shape dummies may be strided)
  tlev_ptr = c_null_ptr
  if (present(tlev)) then
    tlev_c = tlev
    tlev_ptr = c_loc(tlev_c)
  endif
  rc = ecckd_gas_optics_lw_host_f64(this%device_model, ncol, nlay, c_loc(plev_c), c_loc(tlay_c), tlev_ptr, c_loc(tsfc_c), &
         ngas, c_loc(name_ptrs), c_loc(views), c_loc(optical_props%tau), c_loc(sources%lay_source), &
         c_loc(sources%lev_source_inc), c_loc(sources%lev_source_dec), c_loc(sources%sfc_source), 0, err)
  error_msg = ""
  if (rc .ne. 0) call c_to_f_string(err, error_msg)       ! e.g. "tlev is required for ecckd" (reference :415)
end function gas_optics_int


!> @brief Compute gas optical depth, ssa, g and the TOA source (reference :431-473).
function gas_optics_ext(this, play, plev, tlay, gas_desc, optical_props, toa_src, col_dry) &
  result(error_msg)

  class(ty_gas_optics_ecckd), intent(in) :: this
  real(kind=wp), dimension(:,:), intent(in) :: play
  real(kind=wp), dimension(:,:), intent(in) :: plev
  real(kind=wp), dimension(:,:), intent(in) :: tlay
  type(ty_gas_concs), intent(in) :: gas_desc
  class(ty_optical_props_arry), intent(inout) :: optical_props
  real(kind=wp), dimension(:,:), intent(out), target :: toa_src
  real(kind=wp), dimension(:,:), intent(in), target, optional :: col_dry
  character(len=128) :: error_msg

  character(kind=c_char, len=33), dimension(:), allocatable, target :: names
  type(c_ptr), dimension(:), allocatable, target :: name_ptrs
  type(ecckd_vmr_view), dimension(:), allocatable, target :: views
  real(kind=wp), dimension(:,:,:), allocatable, target :: vmr
  real(kind=wp), dimension(:,:), allocatable, target :: plev_c, tlay_c, toa_c
  character(kind=c_char), dimension(128) :: err
  type(c_ptr) :: ssa_ptr, g_ptr
  integer :: ncol, nlay, ngas, rc

  if (.not. c_associated(this%device_model)) then
    error_msg = "ecckd: call finalize_load() after load_and_init (device tables not packed)"
    return
  endif
  ncol = size(tlay, 1); nlay = size(tlay, 2)
  error_msg = marshal_gases(gas_desc, ncol, nlay, names, name_ptrs, views, vmr, ngas)
  if (len_trim(error_msg) .gt. 0) return
  plev_c = plev; tlay_c = tlay
  allocate(toa_c(ncol, size(toa_src, 2)))
  ssa_ptr = c_null_ptr; g_ptr = c_null_ptr
  select type(optical_props)
    type is (ty_optical_props_2str)
      ssa_ptr = c_loc(optical_props%ssa); g_ptr = c_loc(optical_props%g)
  end select                                              ! not 2str: the library returns the reference's error (:462)
  rc = ecckd_gas_optics_sw_host_f64(this%device_model, ncol, nlay, c_loc(plev_c), c_loc(tlay_c), ngas, c_loc(name_ptrs), &
         c_loc(views), c_loc(optical_props%tau), ssa_ptr, g_ptr, c_loc(toa_c), 0, err)
  error_msg = ""
  if (rc .ne. 0) then
    call c_to_f_string(err, error_msg)
    return
  endif
  toa_src(:,:) = toa_c(:,:)
end function gas_optics_ext


subroutine c_to_f_string(c, f)
  character(kind=c_char), dimension(:), intent(in) :: c
  character(len=*), intent(out) :: f
  integer :: i
  f = ""
  do i = 1, min(size(c), len(f))
    if (c(i) .eq. c_null_char) exit
    f(i:i) = c(i)
  enddo
end subroutine c_to_f_string


! ---- getters: reference :477-553, unchanged ------------------------------------------------
pure function get_ngas(this)
  class(ty_gas_optics_ecckd), intent(in) :: this
  integer :: get_ngas
  get_ngas = this%num_gases
end function get_ngas

pure function source_is_internal(this)
  class(ty_gas_optics_ecckd), intent(in) :: this
  logical :: source_is_internal
  source_is_internal = allocated(this%temperature_planck)
end function source_is_internal

pure function source_is_external(this)
  class(ty_gas_optics_ecckd), intent(in) :: this
  logical :: source_is_external
  source_is_external = allocated(this%solar_irradiance)
end function source_is_external

pure function get_gases(this)
  class(ty_gas_optics_ecckd), intent(in) :: this
  character(len=32), dimension(this%num_gases) :: get_gases
  get_gases = this%gas(1:this%num_gases)
end function get_gases

pure function get_press_min(this)
  class(ty_gas_optics_ecckd), intent(in) :: this
  real(kind=wp) :: get_press_min
  get_press_min = exp(this%log_pressure(1))
end function get_press_min

pure function get_press_max(this)
  class(ty_gas_optics_ecckd), intent(in) :: this
  real(kind=wp) :: get_press_max
  get_press_max = exp(this%log_pressure(size(this%log_pressure)))
end function get_press_max

pure function get_temp_min(this)
  class(ty_gas_optics_ecckd), intent(in) :: this
  real(kind=wp) :: get_temp_min
  get_temp_min = minval(this%temperature)
end function get_temp_min

pure function get_temp_max(this)
  class(ty_gas_optics_ecckd), intent(in) :: this
  real(kind=wp) :: get_temp_max
  get_temp_max = maxval(this%temperature)
end function get_temp_max


end module gas_optics_ecckd
