!> @file gas_optics_ecckd_b200_shim.f90
!!
!! Drop-in replacement of /root/reference/src/gas_optics_ecckd.f90 for the gas-optics hot
!! path: the module name, the public type ty_gas_optics_ecckd with its public components, the
!! AbsorptionTable helper type, the concentration-dependence codes and every type-bound
!! procedure keep the reference's names, argument lists, dummy characteristics and error
!! strings, so the reference's own loader (example/rfmip-rad-irf/mo_load_coefficients.F90) and
!! drivers (ecckd_rfmip_lw.F90:78-120, ecckd_rfmip_sw.F90:74-118) compile and run against it
!! UNCHANGED: load_and_init, then gas_optics, nothing in between.  What changes is the body of
!! gas_optics_int / gas_optics_ext: instead of the Fortran loop nests (:64-241, :245-289,
!! :293-319 of the reference) they hand the caller's arrays to the C ABI of libecckd_b200.so
!! (include/ecckd_b200.h), which runs sm_100a CUDA kernels.  There is no CPU fallback: if the
!! library or a device is missing the error string says so and the driver's stop_on_err aborts.
!!
!! Device model: `this` is intent(in) in both entry points, so the opaque ecckd_model handle
!! lives in a module-level cache keyed on a signature of the tables; it is created on the first
!! gas_optics call of an object (tables packed once, then device-resident) and shared by objects
!! holding the same tables.  finalize_load() does the same eagerly; release_device() frees it.
!!
!! Array arguments go to bind(C) interfaces with assumed-size dummies (sequence association:
!! contiguous actuals are passed in place, others through the compiler's copy-in/copy-out), so no
!! c_loc is ever taken of a dummy that lacks TARGET; optional arrays (tlev; ssa/g of a
!! one-stream props object) map to NULL through optional bind(C) dummies (TS 29113 / F2018).
!! The generic names ecckd_lw_host / ecckd_sw_host resolve to the _f64 or _f32 entry points by
!! the kind of wp at compile time.
!!
!! NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Fortran compiler.  INTEGRATION.md
!! gives the build line (and names the stub rte-rrtmgp modules it can be compiled against); tests/c/shim_sequence.c
!! replays this file's call sequence (model create from host tables with nband = 0 and blank-padded
!! names, host entry points on pageable memory, error strings) in plain C on the GPU.
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


!> @brief ecckd gas optics: the reference's components (:23-36), nothing added.
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
  procedure, public :: finalize_load    !< optional: pack the tables now instead of at the first call
  procedure, public :: release_device   !< optional: free the device model of these tables
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
end interface

!> host-array entry points, generic over the kind of wp (include/ecckd_b200.h: ecckd_gas_optics_{lw,sw}_host_{f64,f32})
interface ecckd_lw_host
  integer(c_int) function ecckd_gas_optics_lw_host_f64(model, ncol, nlay, plev, tlay, tlev, tsfc, ngas, names, vmr, &
      tau, lay_source, lev_source_inc, lev_source_dec, sfc_source, block_cols, err) bind(c, name="ecckd_gas_optics_lw_host_f64")
    import :: c_int, c_ptr, c_char, c_double, ecckd_vmr_view
    type(c_ptr), value :: model
    integer(c_int), value :: ncol, nlay, ngas, block_cols
    real(c_double), dimension(*), intent(in) :: plev, tlay, tsfc
    real(c_double), dimension(*), intent(in), optional :: tlev
    type(c_ptr), dimension(*), intent(in) :: names
    type(ecckd_vmr_view), dimension(*), intent(in) :: vmr
    real(c_double), dimension(*), intent(inout) :: tau, lay_source, lev_source_inc, lev_source_dec, sfc_source
    character(kind=c_char) :: err(128)
  end function
  integer(c_int) function ecckd_gas_optics_lw_host_f32(model, ncol, nlay, plev, tlay, tlev, tsfc, ngas, names, vmr, &
      tau, lay_source, lev_source_inc, lev_source_dec, sfc_source, block_cols, err) bind(c, name="ecckd_gas_optics_lw_host_f32")
    import :: c_int, c_ptr, c_char, c_float, ecckd_vmr_view
    type(c_ptr), value :: model
    integer(c_int), value :: ncol, nlay, ngas, block_cols
    real(c_float), dimension(*), intent(in) :: plev, tlay, tsfc
    real(c_float), dimension(*), intent(in), optional :: tlev
    type(c_ptr), dimension(*), intent(in) :: names
    type(ecckd_vmr_view), dimension(*), intent(in) :: vmr
    real(c_float), dimension(*), intent(inout) :: tau, lay_source, lev_source_inc, lev_source_dec, sfc_source
    character(kind=c_char) :: err(128)
  end function
end interface ecckd_lw_host

interface ecckd_sw_host
  integer(c_int) function ecckd_gas_optics_sw_host_f64(model, ncol, nlay, plev, tlay, ngas, names, vmr, &
      tau, ssa, g, toa_src, block_cols, err) bind(c, name="ecckd_gas_optics_sw_host_f64")
    import :: c_int, c_ptr, c_char, c_double, ecckd_vmr_view
    type(c_ptr), value :: model
    integer(c_int), value :: ncol, nlay, ngas, block_cols
    real(c_double), dimension(*), intent(in) :: plev, tlay
    type(c_ptr), dimension(*), intent(in) :: names
    type(ecckd_vmr_view), dimension(*), intent(in) :: vmr
    real(c_double), dimension(*), intent(inout) :: tau, toa_src
    real(c_double), dimension(*), intent(inout), optional :: ssa, g
    character(kind=c_char) :: err(128)
  end function
  integer(c_int) function ecckd_gas_optics_sw_host_f32(model, ncol, nlay, plev, tlay, ngas, names, vmr, &
      tau, ssa, g, toa_src, block_cols, err) bind(c, name="ecckd_gas_optics_sw_host_f32")
    import :: c_int, c_ptr, c_char, c_float, ecckd_vmr_view
    type(c_ptr), value :: model
    integer(c_int), value :: ncol, nlay, ngas, block_cols
    real(c_float), dimension(*), intent(in) :: plev, tlay
    type(c_ptr), dimension(*), intent(in) :: names
    type(ecckd_vmr_view), dimension(*), intent(in) :: vmr
    real(c_float), dimension(*), intent(inout) :: tau, toa_src
    real(c_float), dimension(*), intent(inout), optional :: ssa, g
    character(kind=c_char) :: err(128)
  end function
end interface ecckd_sw_host


! ---- device models, keyed on a signature of the tables (this is intent(in) in gas_optics) ---
integer, parameter :: max_models = 16
integer, parameter :: nsig = 6 + 3 * 16
type :: cached_model
  type(c_ptr) :: handle = c_null_ptr
  real(kind=wp), dimension(nsig) :: sig = 0._wp
end type cached_model
type(cached_model), dimension(max_models), save :: cache
integer, save :: ncached = 0


contains


!> @brief O(num_gases) signature of the tables: extents, grid ends and a few entries of every table.
pure function table_signature(this) result(sig)
  class(ty_gas_optics_ecckd), intent(in) :: this
  real(kind=wp), dimension(nsig) :: sig
  integer :: i, n1, n2, n3, n4

  sig(:) = 0._wp
  sig(1) = merge(1._wp, 0._wp, this%shortwave)
  sig(2) = real(size(this%gpoint_fraction, 2), wp)
  sig(3) = real(size(this%log_pressure), wp) + 1000._wp * real(size(this%temperature, 2), wp)
  sig(4) = real(this%num_gases, wp)
  sig(5) = this%log_pressure(1) + this%log_pressure(size(this%log_pressure))
  sig(6) = this%temperature(1, 1) + this%temperature(size(this%temperature, 1), size(this%temperature, 2))
  do i = 1, this%num_gases
    n1 = size(this%absorption(i)%coefficient, 1); n2 = size(this%absorption(i)%coefficient, 2)
    n3 = size(this%absorption(i)%coefficient, 3); n4 = size(this%absorption(i)%coefficient, 4)
    sig(6 + 3 * i - 2) = this%absorption(i)%coefficient(1, 1, 1, 1) + this%absorption(i)%coefficient(n1, n2, n3, n4)
    sig(6 + 3 * i - 1) = this%absorption(i)%coefficient((n1 + 1) / 2, (n2 + 1) / 2, (n3 + 1) / 2, (n4 + 1) / 2)
    sig(6 + 3 * i) = this%absorption(i)%reference_mole_fraction + real(this%absorption(i)%concentration_dependence_code, wp)
  enddo
end function table_signature


!> @brief Handle of the device model of these tables; created (tables packed, uploaded) on first use.
function device_model(this, error_msg) result(handle)
  class(ty_gas_optics_ecckd), intent(in) :: this
  character(len=128), intent(out) :: error_msg
  type(c_ptr) :: handle

  real(kind=wp), dimension(nsig) :: sig
  integer :: i

  error_msg = ""
  handle = c_null_ptr
  sig = table_signature(this)
  !$omp critical (ecckd_b200_model_cache)
  do i = 1, ncached
    if (c_associated(cache(i)%handle) .and. all(cache(i)%sig .eq. sig)) then
      handle = cache(i)%handle
      exit
    endif
  enddo
  if (.not. c_associated(handle)) then
    if (ncached .ge. max_models) then
      error_msg = "ecckd: more than 16 distinct ck models alive (release_device() frees one)"
    else
      handle = create_model(this, error_msg)
      if (c_associated(handle)) then
        ncached = ncached + 1
        cache(ncached)%handle = handle
        cache(ncached)%sig = sig
      endif
    endif
  endif
  !$omp end critical (ecckd_b200_model_cache)
end function device_model


!> @brief Pack the tables the loader filled (mo_load_coefficients.F90:46-144) and create the device model.
!! The tables are copied into local TARGET arrays first: `this` has no TARGET attribute, and the
!! library copies what it is given before returning.
function create_model(this, error_msg) result(handle)
  class(ty_gas_optics_ecckd), intent(in) :: this
  character(len=128), intent(out) :: error_msg
  type(c_ptr) :: handle

  type :: gas_copy
    real(c_double), dimension(:), allocatable :: coefficient, mole_fraction
  end type gas_copy
  type(gas_copy), dimension(16), target :: gc
  real(c_double), dimension(:), allocatable, target :: log_pressure, temperature, temperature_planck, planck_function
  real(c_double), dimension(:), allocatable, target :: rayleigh, solar
  type(ecckd_host_tables) :: t
  type(ecckd_gas_table), dimension(16), target :: g
  character(kind=c_char), dimension(128) :: err
  integer :: i, k, rc

  error_msg = ""
  handle = c_null_ptr
  log_pressure = real(this%log_pressure, c_double)
  temperature = real(reshape(this%temperature, [size(this%temperature)]), c_double)
  t%shortwave = merge(1, 0, this%shortwave)
  t%ngpt = size(this%gpoint_fraction, 2)
  t%npress = size(this%log_pressure)
  t%ntemp = size(this%temperature, 2)
  t%nplanck = 0
  t%num_gases = this%num_gases
  t%log_pressure = c_loc(log_pressure)
  t%temperature = c_loc(temperature)
  t%temperature_planck = c_null_ptr; t%planck_function = c_null_ptr
  t%rayleigh_molar_scattering_coeff = c_null_ptr; t%solar_irradiance = c_null_ptr
  if (allocated(this%temperature_planck)) then
    temperature_planck = real(this%temperature_planck, c_double)
    planck_function = real(reshape(this%planck_function, [size(this%planck_function)]), c_double)
    t%nplanck = size(this%temperature_planck)
    t%temperature_planck = c_loc(temperature_planck)
    t%planck_function = c_loc(planck_function)
  endif
  if (allocated(this%solar_irradiance)) then
    solar = real(this%solar_irradiance, c_double)
    rayleigh = real(this%rayleigh_molar_scattering_coeff, c_double)
    t%solar_irradiance = c_loc(solar)
    t%rayleigh_molar_scattering_coeff = c_loc(rayleigh)
  endif
  do i = 1, this%num_gases
    do k = 1, 32
      g(i)%name(k) = this%gas(i)(k:k)                     ! blank-padded, as the loader leaves it
    enddo
    g(i)%composite_only = merge(1, 0, this%absorption(i)%composite_only)
    g(i)%concentration_dependence_code = this%absorption(i)%concentration_dependence_code
    g(i)%n_mole_fraction = 0
    g(i)%mole_fraction = c_null_ptr
    if (allocated(this%absorption(i)%mole_fraction)) then
      gc(i)%mole_fraction = real(this%absorption(i)%mole_fraction, c_double)
      g(i)%n_mole_fraction = size(this%absorption(i)%mole_fraction)
      g(i)%mole_fraction = c_loc(gc(i)%mole_fraction)
    endif
    g(i)%reserved = 0
    g(i)%reference_mole_fraction = this%absorption(i)%reference_mole_fraction
    gc(i)%coefficient = real(reshape(this%absorption(i)%coefficient, [size(this%absorption(i)%coefficient)]), c_double)
    g(i)%coefficient = c_loc(gc(i)%coefficient)
  enddo
  t%gas = c_loc(g)
  t%nband = 0; t%reserved = 0
  t%band_lims_wvn = c_null_ptr; t%band2gpt = c_null_ptr      ! kept by the inherited ty_optical_props%init
  rc = ecckd_model_create(t, 1, c_null_ptr, handle, err)
  if (rc .ne. 0) then
    call c_to_f_string(err, error_msg)
    handle = c_null_ptr
  endif
end function create_model


!> @brief Optional: create the device model now (idempotent; gas_optics does it on first use otherwise).
function finalize_load(this) result(error_msg)
  class(ty_gas_optics_ecckd), intent(in) :: this
  character(len=128) :: error_msg
  type(c_ptr) :: handle
  handle = device_model(this, error_msg)
end function finalize_load


!> @brief Free the device model of these tables (if any).
subroutine release_device(this)
  class(ty_gas_optics_ecckd), intent(in) :: this
  real(kind=wp), dimension(nsig) :: sig
  integer :: i
  sig = table_signature(this)
  !$omp critical (ecckd_b200_model_cache)
  do i = 1, ncached
    if (c_associated(cache(i)%handle) .and. all(cache(i)%sig .eq. sig)) then
      call ecckd_model_destroy(cache(i)%handle)
      cache(i) = cache(ncached)
      cache(ncached)%handle = c_null_ptr
      ncached = ncached - 1
      exit
    endif
  enddo
  !$omp end critical (ecckd_b200_model_cache)
end subroutine release_device


!> @brief Gather the caller's ty_gas_concs into names[] + ecckd_vmr_view[] (reference :340-357).
!! get_vmr materialises (ncol,nlay) exactly as the reference does (:343,351); a field that is
!! constant is handed over as a scalar view so that the library can pre-mix it (its fast path).
function marshal_gases(gas_desc, ncol, nlay, names, name_ptrs, views, vmr, ngas) result(error_msg)
  type(ty_gas_concs), intent(in) :: gas_desc
  integer, intent(in) :: ncol, nlay
  character(kind=c_char), dimension(:,:), allocatable, target, intent(out) :: names   ! (33, ngas), NUL-terminated
  type(c_ptr), dimension(:), allocatable, intent(out) :: name_ptrs
  type(ecckd_vmr_view), dimension(:), allocatable, intent(out) :: views
  real(kind=wp), dimension(:,:,:), allocatable, target, intent(out) :: vmr
  integer, intent(out) :: ngas
  character(len=128) :: error_msg

  character(len=32), dimension(:), allocatable :: gas_names
  integer :: j, k, n

  error_msg = ""
  ngas = gas_desc%get_num_gases()
  allocate(gas_names(ngas), names(33, ngas), name_ptrs(ngas), views(ngas), vmr(ncol, nlay, ngas))
  gas_names = gas_desc%get_gas_names()
  do j = 1, ngas
    n = len_trim(gas_names(j))
    do k = 1, 33
      names(k, j) = c_null_char
      if (k .le. n) names(k, j) = gas_names(j)(k:k)
    enddo
    name_ptrs(j) = c_loc(names(1, j))
    error_msg = gas_desc%get_vmr(gas_names(j), vmr(:,:,j))
    if (trim(error_msg) .ne. "") return
    views(j)%ptr = c_loc(vmr(1, 1, j))
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

  character(kind=c_char), dimension(:,:), allocatable, target :: names
  type(c_ptr), dimension(:), allocatable :: name_ptrs
  type(ecckd_vmr_view), dimension(:), allocatable :: views
  real(kind=wp), dimension(:,:,:), allocatable, target :: vmr
  character(kind=c_char), dimension(128) :: err
  type(c_ptr) :: model
  integer :: ncol, nlay, ngas, rc

  model = device_model(this, error_msg)
  if (.not. c_associated(model)) return
  ncol = size(tlay, 1); nlay = size(tlay, 2)
  error_msg = marshal_gases(gas_desc, ncol, nlay, names, name_ptrs, views, vmr, ngas)
  if (len_trim(error_msg) .gt. 0) return
  ! tlev is forwarded as it is: an absent optional becomes NULL, and the library then returns the
  ! reference's "tlev is required for ecckd" (:415) after writing tau, lay_source and sfc_source
  rc = ecckd_lw_host(model, ncol, nlay, plev, tlay, tlev, tsfc, ngas, name_ptrs, views, optical_props%tau, &
                     sources%lay_source, sources%lev_source_inc, sources%lev_source_dec, sources%sfc_source, 0, err)
  error_msg = ""
  if (rc .ne. 0) call c_to_f_string(err, error_msg)
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
  real(kind=wp), dimension(:,:), intent(out) :: toa_src
  real(kind=wp), dimension(:,:), intent(in), target, optional :: col_dry
  character(len=128) :: error_msg

  character(kind=c_char), dimension(:,:), allocatable, target :: names
  type(c_ptr), dimension(:), allocatable :: name_ptrs
  type(ecckd_vmr_view), dimension(:), allocatable :: views
  real(kind=wp), dimension(:,:,:), allocatable, target :: vmr
  character(kind=c_char), dimension(128) :: err
  type(c_ptr) :: model
  integer :: ncol, nlay, ngas, rc

  model = device_model(this, error_msg)
  if (.not. c_associated(model)) return
  ncol = size(tlay, 1); nlay = size(tlay, 2)
  error_msg = marshal_gases(gas_desc, ncol, nlay, names, name_ptrs, views, vmr, ngas)
  if (len_trim(error_msg) .gt. 0) return
  select type(optical_props)
    type is (ty_optical_props_2str)
      rc = ecckd_sw_host(model, ncol, nlay, plev, tlay, ngas, name_ptrs, views, optical_props%tau, &
                         optical_props%ssa, optical_props%g, toa_src, 0, err)
    class default      ! ssa, g absent -> NULL: tau is written, then the reference's error (:462) comes back
      rc = ecckd_sw_host(model, ncol, nlay, plev, tlay, ngas, name_ptrs, views, optical_props%tau, &
                         toa_src=toa_src, block_cols=0, err=err)
  end select
  error_msg = ""
  if (rc .ne. 0) call c_to_f_string(err, error_msg)
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
